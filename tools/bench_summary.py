"""Print the headline and every workload record of a bench.py JSON line in a few lines each."""
import json
import sys


def one(name, d):
    r = d.get("roofline") or {}
    rs = d.get("roofline_step") or {}
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print("%-8s %-24s %.4g %s  ms/step %.4g  roof %s=%.3f (%s)  step-roof %.3f  e2e ms %.4g (%.4g)  cpu %.4g %s x%s  wall %.1fs"
          % (name, d.get("metric"), d["value"], d.get("unit", ""), d["ms_per_step"], r.get("kernel"), r.get("frac") or -1,
             r.get("bound"), rs.get("frac") or -1, e.get("ms_per_step") or -1, e.get("value") or -1, c.get("value") or -1,
             c.get("kind"), c.get("cores"), d.get("bench_wall_s") or -1))
    if c.get("variants"):
        print("         cpu variants:", {k: "%.4g" % v["value"] for k, v in c["variants"].items()})
    print("         check:", d.get("check"), " launches/step:", d.get("gpu_launches"), " clocks:", d.get("clocks"))
    if d.get("modes"):
        print("         modes:", {k: round(v["ms_per_step"], 3) for k, v in d["modes"].items()})
    if d.get("kernels"):
        print("         kernels:", {k: (round(v["avg_us"], 1), round(v["launches_per_step"], 2), round(v.get("frac") or 0, 3)) for k, v in d["kernels"].items()})
    if d.get("branch_loop"):
        b = d["branch_loop"]
        print("         branch loop: total %.3g ms, eval %.3g ms, reprune %.3g ms" % (b["loop_total_ms"], b["lnl_d1_d2_eval_ms"], b["reprune_ms"]))
    if d.get("roofline_tensor"):
        t = d["roofline_tensor"]
        print("         tensor roofline: %s useful %.3g TF (%.3f of the fp64 DMMA peak), issued %s" % (
            t.get("kernel"), t.get("achieved") or -1, t.get("frac") or -1, t.get("issued_frac")))
    if d.get("directions"):
        print("         directions:", d["directions"])
    if d.get("group"):
        print("         group:", d["group"])


def main(path):
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    if not lines:
        print("no JSON line in", path)
        return
    d = json.loads(lines[-1])
    one("HEAD", d)
    for k, v in (d.get("workloads") or {}).items():
        one(k, v)


if __name__ == "__main__":
    main(sys.argv[1])
