"""Harness mirror of the reference's built-in alphabets (lib/alphabet.ml) as 256-entry symbol ->
state-mask tables for phylo_engine_set_symbol_table. Test / bench support only: in phylocaml
itself the table is read off Alphabet.t's name_code map.

BitFlag alphabets number their states 1, 2, 4, ... in list order and an equate is the OR of its
states (lib/alphabet.ml:180-183, :196-199); names are upper-cased when `case:false` (:185), so
both cases of a letter map to the same code here."""
import numpy as np

GAP, MISSING = "-", "?"  # default_gap, default_missing (lib/alphabet.ml:110,114)


def _bitflag(states, equates):
    code = {s: 1 << i for i, s in enumerate(states)}
    for name, members in equates:
        v = 0
        for m in members:
            v |= code[m]
        code[name] = v
    return code


def _table(code):
    t = np.zeros(256, dtype=np.uint64)
    for name, v in code.items():
        if len(name) != 1:
            continue
        t[ord(name)] = v
        if name.isalpha():
            t[ord(name.lower())] = v
            t[ord(name.upper())] = v
    return t


def dna_codes():
    """Alphabet.dna (lib/alphabet.ml:301-307): A C G T - X = 1 2 4 8 16 32; equates 0..4 -> A C G T -.
    test/alphabetTest.ml:12-18 pins 1,2,4,8,16 -> A,C,G,T,-."""
    return _bitflag(["A", "C", "G", "T", GAP, "X"],
                    [("0", ["A"]), ("1", ["C"]), ("2", ["G"]), ("3", ["T"]), ("4", [GAP])])


def nucleotides_codes():
    """Alphabet.nucleotides (lib/alphabet.ml:309-326): A C G T - = 1 2 4 8 16, IUPAC and
    indel-polymorphism letters as equates, ? = all five."""
    return _bitflag(["A", "C", "G", "T", GAP], [
        ("M", ["A", "C"]), ("R", ["A", "G"]), ("W", ["A", "T"]), ("S", ["G", "C"]), ("Y", ["T", "C"]),
        ("K", ["G", "T"]), ("V", ["G", "T", "C"]), ("H", ["G", "T", "A"]), ("D", ["C", "T", "A"]),
        ("B", ["G", "C", "A"]), ("N", ["A", "C", "G", "T"]), ("X", ["A", "C", "G", "T"]),
        ("1", ["T", GAP]), ("2", ["G", GAP]), ("3", ["G", "T", GAP]), ("4", ["C", GAP]), ("5", ["T", "C", GAP]),
        ("6", ["G", "C", GAP]), ("7", ["G", "T", "C", GAP]), ("8", ["A", GAP]), ("9", ["T", "A", GAP]),
        ("0", ["G", "A", GAP]), ("E", ["G", "T", "A", GAP]), ("F", ["A", "C", GAP]), ("I", ["T", "A", "C", GAP]),
        ("J", ["G", "A", "C", GAP]), ("P", ["G", "T", "A", "C", GAP]), (MISSING, ["G", "T", "A", "C", GAP])])


AMINOACIDS = list("ARNDCQEGHILKMFPSTWYV")  # the 20 residues in lib/alphabet.ml:330-339 order; then X, -


def dna_table():
    return _table(dna_codes())


def nucleotides_table():
    """Parsimony view: gap is a fifth state."""
    return _table(nucleotides_codes())


def nucleotides_table_likelihood():
    """4-state likelihood view of Alphabet.nucleotides: the gap bit is dropped and a cell that is
    only gap (or missing) stands for all four states (`Missing`, lib/mlModel.mli:75-76)."""
    t = nucleotides_table()
    known = t != 0
    t = t & np.uint64(15)
    t[known & (t == 0)] = 15
    return t


def aminoacids_table_likelihood():
    """Alphabet.aminoacids is Sequential (codes 0..21, lib/alphabet.ml:328-343); likelihood wants
    one bit per residue: residue i -> 1 << i, X and - -> all twenty."""
    code = {a: 1 << i for i, a in enumerate(AMINOACIDS)}
    code["X"] = code[GAP] = (1 << 20) - 1
    return _table(code)


def translate(table, symbols):
    """CPU statement of the translation (numpy gather) for tests."""
    return table[np.asarray(symbols, dtype=np.uint8)]
