"""Shared parity-case catalogue: (name, fastq bytes, dna_order, quality_order, plus_rep).
dna_order = 3 * CLI -d level, quality_order = CLI -q level (src/DsrcOperator.h:74-90)."""
import os

import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_cases():
    c = []
    ill_b = synth.illumina(600, seed=11, regime="binned")
    ill_f = synth.illumina(500, seed=12, regime="full")
    ion = synth.ion454(250, seed=5)
    for d, q in [(0, 0), (3, 1), (6, 2), (9, 2), (0, 2), (6, 0), (3, 2), (9, 1)]:
        c.append(("ill_binned_d%d_q%d" % (d, q), ill_b, d, q, 0))
        c.append(("ill_full_d%d_q%d" % (d, q), ill_f, d, q, 0))
    for d, q in [(3, 1), (6, 2), (9, 2), (9, 1)]:
        c.append(("ion454_d%d_q%d" % (d, q), ion, d, q, 0))
    ion_plain = synth.ion454(200, seed=6, iupac=False)
    c.append(("ion454_plain_d0_q0", ion_plain, 0, 0, 0))
    c.append(("ill_barcode_d6_q2", synth.illumina(400, seed=13, barcode_var=True), 6, 2, 0))
    c.append(("ill_barcode_d0_q0", synth.illumina(400, seed=13, barcode_var=True), 0, 0, 0))
    c.append(("ill_plusrep_d6_q2", synth.illumina(300, seed=14, plus_rep=True), 6, 2, 1))
    c.append(("mixed_titles_d6_q2", synth.mixed_titles(300), 6, 2, 0))
    c.append(("mixed_titles_d0_q0", synth.mixed_titles(300), 0, 0, 0))
    c.append(("ill_notail_d0_q0", synth.illumina(300, seed=15, tail_frac=0.0, regime="full"), 0, 0, 0))
    c.append(("ill_alltail_d0_q0", synth.illumina(300, seed=16, tail_frac=0.9, regime="full"), 0, 0, 0))
    for n in (1, 2, 3, 5):
        c.append(("tiny%d_d6_q2" % n, synth.illumina(n, seed=20 + n, regime="full"), 6, 2, 0))
        c.append(("tiny%d_d0_q0" % n, synth.illumina(n, seed=20 + n, regime="full"), 0, 0, 0))
    c.append(("ill_long_d6_q2", synth.illumina(3000, seed=17), 6, 2, 0))
    c.append(("ill_startbig_d6_q2", synth.illumina(700, seed=18, start_index=99990), 6, 2, 0))
    c.append(("q1_valuevar_d6_q2", synth.illumina(40, seed=2, small_field=True), 6, 2, 0))
    c.append(("q1_valuevar_big_d0_q0", synth.illumina(900, seed=3, small_field=True), 0, 0, 0))
    sk = synth.skewed(5000)
    for d, q in [(6, 2), (9, 1), (3, 2)]:
        c.append(("skewed_rescale_d%d_q%d" % (d, q), sk, d, q, 0))
    rq = synth.random_quals(1500)
    for d, q in [(6, 2), (9, 1)]:
        c.append(("randq60_d%d_q%d" % (d, q), rq, d, q, 0))
    # the 16-symbol quality models by symbol count (the GPU scan engine is compiled per 4 / 8 / 16 live symbols): sticky and random
    m12 = synth.illumina(900, seed=31, regime="mid12")
    two = synth.illumina(900, seed=32, regime="two")
    for d, q in [(6, 2), (3, 1)]:
        c.append(("ill_mid12_d%d_q%d" % (d, q), m12, d, q, 0))
        c.append(("ill_two_d%d_q%d" % (d, q), two, d, q, 0))
    c.append(("randq12_d6_q2", synth.random_quals(900, seed=33, n_levels=12), 6, 2, 0))
    c.append(("randq7_d6_q2", synth.random_quals(900, seed=34, n_levels=7), 6, 2, 0))
    c.append(("randq3_d9_q1", synth.random_quals(900, seed=35, n_levels=3), 9, 1, 0))
    # the same models with a different read length per record (position bucket per record instead of the fixed-length table)
    vl4 = synth.varlen_binned(700, seed=41)
    vl12 = synth.varlen_binned(700, seed=42, levels=(2, 6, 9, 12, 15, 18, 21, 24, 27, 30, 34, 38))
    c.append(("varlen_binned_d6_q2", vl4, 6, 2, 0))
    c.append(("varlen_binned_d3_q1", vl4, 3, 1, 0))
    c.append(("varlen_mid12_d6_q2", vl12, 6, 2, 0))
    # inputs found by tools/fuzz_parity.py: a two-symbol Huffman tree whose zero-frequency symbol is the larger one (the reference
    # forces both frequencies to 1 in place, src/huffman.cpp:128-133, so the former minimum stays the left child), and very short
    # variable-length reads with '+' title repetition
    c.append(("fuzz_huffman_two_symbol_tie_d3_q0", open(os.path.join(GOLDEN_DIR, "fuzz_huffman_two_symbol_tie.fq"), "rb").read(), 3, 0, 1))
    # read IDs with two 100-160 character text fields: > 256 per-position Huffman trees in the tag header, the compressed block is
    # larger than the input
    lt = open(os.path.join(GOLDEN_DIR, "fuzz_long_text_fields.fq"), "rb").read()
    c.append(("fuzz_long_text_fields_d6_q2", lt, 6, 2, 0))
    c.append(("fuzz_long_text_fields_d0_q0", lt, 0, 0, 0))
    c.append(("fuzz_short_reads_d6_q2", open(os.path.join(GOLDEN_DIR, "fuzz_short_reads.fq"), "rb").read(), 6, 2, 0))
    return c
