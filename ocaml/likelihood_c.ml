(* likelihood_c.ml -- body for the reference's empty lib/likelihood_c.ml (all
   `failwith "TODO"`, lib/likelihood_c.ml:1-33), written to the unchanged NodeData.S
   signature (lib/nodeData.ml:3-35) over the B200 engine's C stubs (stubs/phylo_stubs.c).

   NOT COMPILED IN THIS REPOSITORY: the build image has no OCaml toolchain (see DESIGN.md). The
   stubs underneath are compiled and driven from C exactly as this file drives them
   (tests/c/stub_lifetime.c: 1000 successive trees through median_2 on a 2*T-slot engine).

   A node value [t] names a device-resident CLV through a [node] custom block whose finalizer
   returns the slot to the engine (the pattern of the reference's bitvector blocks,
   lib/bitvector/bv.c:183-189,229-244): results are fresh values, inputs are never mutated, and
   the GC decides when device memory is reused. The character set's [spec] (engine + sizes)
   travels inside every [t]; there is no global state, so a node may carry any number of
   likelihood character sets (Node.n.prelim : nd IntMap.t, lib/node.ml:64-74), one engine each. *)
open Internal

type engine
type node                       (* custom block: (engine, slot), finalizer = phylo_lk_node_release *)
type matrix = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array2.t
type vector = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type masks  = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t
type ids    = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array2.t

external engine_create : int -> engine = "phylo_CAML_engine_create"
external custom_max : int -> unit = "phylo_CAML_custom_max"      (* cf. bv_CAML_custom_max, bv.c:250-256 *)
(* PHYLO_OPT_* of include/phylo_engine.h: 2 = retain interior CLVs / state sets (1) or score only (0) *)
external engine_set_option : engine -> int -> int -> unit = "phylo_CAML_engine_set_option"
external node_slot : node -> int = "phylo_CAML_node_slot"
external node_stats : engine -> bool -> int * int * int = "phylo_CAML_node_stats"
external set_model_ : engine -> matrix -> matrix -> matrix option
                      -> (vector * vector * vector * float option) -> unit
                    = "likelihood_CAML_set_model"
external set_tips_ : engine -> masks -> vector option -> int -> unit = "likelihood_CAML_set_tips"
external tip_ : engine -> int -> node = "likelihood_CAML_tip"
external median2_ : engine -> node -> float -> node -> float -> node = "likelihood_CAML_median2"
external median3_ : engine -> (node * float) -> (node * float) -> (node * float) -> node
                  = "likelihood_CAML_median3"
external edge_lnl_ : engine -> node -> node -> vector -> vector -> unit = "likelihood_CAML_edge_lnl"
external optimize_branch_ : engine -> node -> node -> (float * float * float * float) -> float * float
                           = "likelihood_CAML_optimize_branch"
external edge_eval_ : engine -> node -> node -> vector -> matrix -> unit = "likelihood_CAML_edge_eval"
external score_tree_ : engine -> ids -> matrix -> (int * int * float) -> float * node array
                     = "likelihood_CAML_score_tree"

external uppass_ : engine -> ids -> matrix -> (int * int * float) -> node array -> node array
                 = "likelihood_CAML_uppass"
external param_gradient_ : engine -> ids -> matrix -> (int * int * float) -> (node array * node array)
                           -> (matrix option * matrix option * matrix option) -> vector -> unit
                         = "likelihood_CAML_param_gradient_bc" "likelihood_CAML_param_gradient"

type m = MlModel.t

type spec = { engine : engine; n_taxa : int; n_patterns : int; }

(* a directed node: its CLV and the length of the branch above it *)
type t = {
  spec   : spec;
  node   : node;
  branch : float;
  codes  : IntSet.t;
  lnl    : float option;   (* set on root nodes *)
}

let load_model (e : engine) (model : m) =
  set_model_ e model.MlModel.u model.MlModel.d model.MlModel.ui
    (model.MlModel.priors, model.MlModel.rates, model.MlModel.probs, model.MlModel.pinvar)

(* bookkeeping over character codes *)
let filter_codes set t =
  let c = IntSet.inter set t.codes in
  if IntSet.is_empty c then None else Some { t with codes = c }
let filter_codes_comp set t =
  let c = IntSet.diff t.codes set in
  if IntSet.is_empty c then None else Some { t with codes = c }
let cardinal t = t.spec.n_patterns
let get_codes t = t.codes
let mem codes t = match codes with
  | None -> true
  | Some xs -> List.exists (fun x -> IntSet.mem x t.codes) xs
(* a CLV has no set union; the union of two likelihood nodes under parent p is p's own data with the
   children's codes (what Node needs from it: the code bookkeeping of lib/node.ml:95-99) *)
let union p a b = { p with codes = IntSet.union p.codes (IntSet.union a.codes b.codes) }
let compare a b = Pervasives.compare (node_slot a.node) (node_slot b.node)
let recode f t = { t with codes = IntSet.fold (fun x acc -> IntSet.add (f x) acc) t.codes IntSet.empty }

(* kernels *)
let median_1 _ _ x = x
let median_2 _model _prev a b =
  let n = median2_ a.spec.engine a.node a.branch b.node b.branch in
  { a with node = n; lnl = None; codes = IntSet.union a.codes b.codes }
(* the node's CLV conditioned on all three neighbours: (P_a L_a) o (P_b L_b) o (P_c L_c) *)
let median_3 _model _prev a b c =
  let n = median3_ a.spec.engine (a.node, a.branch) (b.node, b.branch) (c.node, c.branch) in
  { a with node = n; lnl = None; codes = IntSet.union a.codes (IntSet.union b.codes c.codes) }
let median_n model prev a = function
  | [b] -> median_2 model prev a b
  | [b; c] -> median_3 model prev a b c
  | _ -> failwith "Likelihood_c.median_n: binary trees only"

(* readjust node n (= median of a and b) against its third neighbour c: the branch above n
   takes its maximum-likelihood length (device sum table + safeguarded Newton). All of n's
   characters count as changed when the length moved. *)
let adjust_3 _model _codes n _a _b c =
  let t, _lnl = optimize_branch_ n.spec.engine n.node c.node (n.branch, 1e-8, 100.0, 1e-8) in
  if abs_float (t -. n.branch) <= 1e-8 *. n.branch then n, IntSet.empty
  else { n with branch = t; lnl = None }, n.codes
let adjust_n model codes n = function
  | [a; b; c] -> adjust_3 model codes n a b c
  | _ -> n, IntSet.empty

(* root edge (a,b) with length t: -lnL (a cost to minimise) *)
let distance_1 _model a b =
  let ts = Bigarray.Array1.of_array Bigarray.float64 Bigarray.c_layout [| a.branch |] in
  let out = Bigarray.Array1.create Bigarray.float64 Bigarray.c_layout 1 in
  edge_lnl_ a.spec.engine a.node b.node ts out;
  ~-. (out.{0})
(* a joined to the edge (b, c): the cost of the tree that results, -lnL of median_3 at its root *)
let distance_2 model a b c = distance_1 model a (median_2 model None b c)
let cost t = match t.lnl with Some l -> ~-. l | None -> 0.0
let root_cost = cost
let leaf_cost _ = 0.0
let to_string t = Printf.sprintf "<clv slot %d, branch %f>" (node_slot t.node) t.branch

(* not part of NodeData.S *)
let of_string _ = failwith "Likelihood_c.of_string: use of_parser"
let of_parser _ = failwith "Likelihood_c.of_parser: data arrive through create_spec"

(* one engine per character set; capacity = 2 * n_taxa slots to start with (the table grows only if
   more node values than that are alive at once) *)
let create_spec ?(device = 0) (model : m) (tips : masks) (weights : vector option) =
  let e = engine_create device in
  let n_taxa = Bigarray.Array2.dim1 tips in
  load_model e model;
  set_tips_ e tips weights (2 * n_taxa);
  { engine = e; n_taxa; n_patterns = Bigarray.Array2.dim2 tips }

(* the node value of taxon i with the branch above it *)
let leaf (s : spec) (i : int) (code : int) (branch : float) =
  { spec = s; node = tip_ s.engine i; branch; codes = IntSet.singleton code; lnl = None }

(* A scorer that wants one launch sequence per tree flattens Tree.post_order_edges
   (lib/tree.ml:171-187) into ids (n_ops x 3: parent, left, right; ids < n_taxa are taxa, the others
   any numbering of the interior nodes) and lens (n_ops x 2) and gets lnL plus the node values. *)
let score_tree (s : spec) (ops : ids) (lens : matrix) root_a root_b root_t =
  score_tree_ s.engine ops lens (root_a, root_b, root_t)

(* ---- alignments as text, and all GPUs of the box ------------------------------------------
   [symbol_table] turns an Alphabet.t into the 256-entry table of state sets the engine applies on
   the device (single-character names only; gap and missing fold into "all states" for likelihood,
   lib/mlModel.mli:75-76). [create_spec_text] is create_spec for an alignment that is still the
   characters of the input file. *)
external set_symbol_table_ :
  engine -> (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array1.t option -> unit
  = "phylo_CAML_set_symbol_table"

let symbol_table (alph : Alphabet.t) (n_states : int) =
  let all = Int64.pred (Int64.shift_left 1L n_states) in
  let t = Bigarray.Array1.create Bigarray.int64 Bigarray.c_layout 256 in
  Bigarray.Array1.fill t 0L;
  StringMap.iter
    (fun name code ->
      if String.length name = 1 then begin
        let m = Int64.logand (Int64.of_int code) all in
        let m = if m = 0L then all else m in          (* gap-only / missing: every state *)
        t.{Char.code (Char.uppercase name.[0])} <- m;
        t.{Char.code (Char.lowercase name.[0])} <- m
      end)
    alph.Alphabet.name_code;
  t

let create_spec_text ?(device = 0) (model : m) (alph : Alphabet.t) (text : masks) (weights : vector option) =
  let e = engine_create device in
  let n_taxa = Bigarray.Array2.dim1 text in
  load_model e model;
  set_symbol_table_ e (Some (symbol_table alph (Bigarray.Array2.dim1 model.MlModel.u)));
  set_tips_ e text weights (2 * n_taxa);
  { engine = e; n_taxa; n_patterns = Bigarray.Array2.dim2 text }


(* ---- beyond NodeData.S: what Node.Make3D and MlModel's optimisers need from the native side ---- *)

(* Third directions for a whole scored tree (Node.Make3D keeps one value per excluded neighbour,
   lib/node.ml:363-477): [down] is score_tree's node array for the schedule [ids]/[lens]; the result holds,
   for op i, the value of the rest of the tree above its left child at 2 i and above its right child at
   2 i + 1. (child, its up value) is then an ordinary directional pair: root_cost / adjust_3 apply to ANY
   branch, and for a reversible model every branch gives the same lnL. *)
let uppass spec ids lens root (down : node array) : node array = uppass_ spec.engine ids lens root down

(* d lnL / d theta for the parameters a model optimiser moves -- the native half of gen_subst_opt_func /
   gen_rates_opt_func / gen_prior_opt_func (lib/mlModel.ml:822-829). The caller describes each parameter by
   dQ/dtheta (rows of S*S), d rates/dtheta (rows of K), d priors/dtheta (rows of S). *)
let param_gradient spec ids lens root ~down ~up ?dq ?drates ?dpriors (n_params : int) : vector =
  let out = Bigarray.Array1.create Bigarray.float64 Bigarray.c_layout n_params in
  param_gradient_ spec.engine ids lens root (down, up) (dq, drates, dpriors) out;
  out
