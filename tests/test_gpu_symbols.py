"""Alphabet symbols -> state masks on the device (phylo_engine_set_symbol_table): an ASCII
alignment scored through the table must give, bit for bit, what the pre-translated masks give."""
import numpy as np
import pytest

from helpers import aa_model, dna_gtr_g4
from phylocaml_b200 import alphabet, engine, tree

pytestmark = pytest.mark.gpu


@pytest.fixture()
def seng(built):
    e = engine.Engine(0)
    yield e
    e.close()


def _ascii_alignment(T, N, letters, seed):
    rng = np.random.default_rng(seed)
    a = np.frombuffer(letters.encode(), dtype=np.uint8)
    return np.ascontiguousarray(a[rng.integers(0, a.size, (T, N))])


def test_dna_likelihood_from_iupac_text(seng, oracle):
    model = dna_gtr_g4()
    T, N = 12, 5003
    tr = tree.random_tree(T, 5)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    text = _ascii_alignment(T, N, "ACGTacgtRYKMSWBDHVN-?", 1)
    tab = alphabet.nucleotides_table_likelihood()
    masks = alphabet.translate(tab, text).astype(np.uint8)
    seng.lk_set_model(model)
    seng.lk_set_tips(masks, capacity=n_nodes)
    want = seng.lk_score_tree(ops, ra, rb, rt)
    site = seng.lk_get_site_lnl()
    seng.set_symbol_table(tab)
    seng.lk_set_tips(text, capacity=n_nodes)
    assert seng.lk_score_tree(ops, ra, rb, rt) == want
    assert np.array_equal(seng.lk_get_site_lnl(), site)
    assert seng.lk_score_alignment(text, ops, ra, rb, rt, capacity=n_nodes) == want
    ref = oracle.lk_score_tree(model, masks, None, ops, n_nodes, ra, rb, rt)["lnl"]
    assert abs(want - ref) <= 1e-9 * abs(ref)
    bad = text.copy()
    bad[3, 77] = ord("Z")  # not in Alphabet.nucleotides
    with pytest.raises(engine.PhyloError) as ei:
        seng.lk_set_tips(bad, capacity=n_nodes)
    assert ei.value.code == -4
    with pytest.raises(engine.PhyloError):
        seng.lk_set_tips(masks.astype(np.uint32), capacity=n_nodes)  # symbols are bytes
    seng.set_symbol_table(None)
    seng.lk_set_tips(masks, capacity=n_nodes)
    assert seng.lk_score_tree(ops, ra, rb, rt) == want


def test_aminoacid_likelihood_from_text(seng):
    model = aa_model(4)
    T, N = 9, 2100
    tr = tree.random_tree(T, 6)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    text = _ascii_alignment(T, N, "".join(alphabet.AMINOACIDS) + "X-", 2)
    tab = alphabet.aminoacids_table_likelihood()
    masks = alphabet.translate(tab, text).astype(np.uint32)
    seng.lk_set_model(model)
    seng.lk_set_tips(masks, capacity=n_nodes)
    want = seng.lk_score_tree(ops, ra, rb, rt)
    seng.set_symbol_table(tab)
    seng.lk_set_tips(text, capacity=n_nodes)
    assert seng.lk_score_tree(ops, ra, rb, rt) == want


def test_fitch_and_compression_from_text(seng, oracle):
    T, N = 16, 40007
    tr = tree.random_tree(T, 7)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    text = _ascii_alignment(T, N, "ACGT" * 6 + "RYN-?acgt", 3)
    tab = alphabet.nucleotides_table()
    codes = alphabet.translate(tab, text).astype(np.uint8)
    want = oracle.fitch_score_tree(codes, None, ops, n_nodes, ra, rb, want_sets=True)
    seng.set_symbol_table(tab)
    seng.fitch_set_tips(text, 5, capacity=n_nodes)
    assert seng.fitch_score_tree(ops, ra, rb) == want["length"]
    node = int(ops[-1]["parent"])
    assert np.array_equal(seng.fitch_get_states(node), want["prelim"][node])
    assert np.array_equal(seng.fitch_get_states(0), codes[0])  # tips come back as codes, not letters
    seng.fitch_set_states(node, want["prelim"][node])          # of_array takes codes even with a table set
    assert np.array_equal(seng.fitch_get_states(node), want["prelim"][node])
    # compression: columns that differ only in case / equivalent letters merge; output is masks
    short = np.ascontiguousarray(text[:, :3000])
    pats, w, s2p = seng.compress_patterns(short)
    seng.set_symbol_table(None)
    pats2, w2, s2p2 = seng.compress_patterns(np.ascontiguousarray(codes[:, :3000]))
    assert np.array_equal(pats, pats2) and np.array_equal(w, w2) and np.array_equal(s2p, s2p2)
    seng.set_symbol_table(tab)
    bad = short.copy()
    bad[0, 1] = ord("!")
    with pytest.raises(engine.PhyloError) as ei:
        seng.compress_patterns(bad)
    assert ei.value.code == -4
