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
    # the shared-memory walk engines (csrc/model_walk.cuh: <= 5 quality symbols and one read length; 4-symbol DNA at order <= 6):
    # random qualities (every lane of a row in another context, 8 position buckets), read lengths that do not divide into the
    # buckets, hot contexts that rescale inside a bucket, a read length below the engine's minimum (the tile engine takes it)
    c.append(("walk_randq4_d6_q2", synth.random_quals(1500, seed=36, n_levels=4), 6, 2, 0))
    c.append(("walk_randq5_d3_q1", synth.random_quals(1200, seed=37, n_levels=5), 3, 1, 0))
    c.append(("walk_len37_hot_d6_q2", synth.skewed(9000, length=37, seed=118), 6, 2, 0))
    c.append(("walk_len101_d3_q2", synth.skewed(4000, length=101, seed=119, p_major=0.6), 3, 2, 0))
    c.append(("walk_len31_d6_q2", synth.skewed(4000, length=31, seed=120, p_major=0.8), 6, 2, 0))
    # read IDs of other platforms and archives (the tag tokenizer's envelope: field counts, title lengths, text fields)
    for style in ("sra", "ont", "pacbio", "bgi"):
        c.append(("ids_%s_d6_q2" % style, synth.read_id_styles(style), 6, 2, 0))
    c.append(("ids_ont_d0_q0", synth.read_id_styles("ont", seed=52), 0, 0, 0))
    c.append(("walk_homopolymer_d6_q1", synth.skewed(3000, length=150, seed=121, p_major=0.995, levels=(2, 37)), 6, 1, 0))
    return c


def scale_cases():
    """Block-scale cases (each >= 256 KiB: every engine crosses many tiles) for the model shapes the small catalogue does not select --
    VERDICT r1 census: -q2 32S / 32F / 128S / 128F, -q1 32S / 128S, -d0 Huffman DNA, Truncated / Plain positional quality, the
    8-symbol DNA models, variable-length 64-symbol quality. expected = (quality scheme byte, DNA scheme byte) the reference writes
    (src/QualityModelerProxy.h:113-122,231-282, src/DnaModelerProxy.h:102-111,160-170); tests/test_scheme_census.py checks them."""
    c = []
    r24 = synth.block_scale(n_levels=24, seed=101, q_lo=7)
    s24 = synth.block_scale(n_levels=24, sticky=True, seed=102, q_lo=7)
    c.append(("scale_q32S_d6_q2", r24, 6, 2, 0, (1, 0)))
    c.append(("scale_q32S_d3_q1", r24, 3, 1, 0, (1, 0)))
    c.append(("scale_q32F_d6_q2", s24, 6, 2, 0, (5, 0)))
    c.append(("scale_q32F_d9_q1", s24, 9, 1, 0, (1, 0)))       # -q1 has no F variants
    wide = dict(n_levels=38, q_lo=7, low_amb=0.05, low_codes=synth.IUPAC_ALL[:10])    # 38 + 10 x 7 = 108 symbols (> 128 is undefined upstream)
    r128 = synth.block_scale(seed=103, **wide)
    s128 = synth.block_scale(seed=104, sticky=True, **wide)
    c.append(("scale_q128S_d6_q2", r128, 6, 2, 0, (3, 0)))
    c.append(("scale_q128S_d3_q1", r128, 3, 1, 0, (3, 0)))
    c.append(("scale_q128F_d6_q2", s128, 6, 2, 0, (7, 0)))
    r41 = synth.block_scale(n_levels=41, seed=105)
    s41 = synth.block_scale(n_levels=41, sticky=True, seed=106)
    c.append(("scale_q64S_d6_q2", r41, 6, 2, 0, (2, 0)))
    c.append(("scale_q64F_d6_q2", s41, 6, 2, 0, (6, 0)))
    c.append(("scale_q64S_d9_q1", s41, 9, 1, 0, (2, 0)))
    v41 = synth.block_scale(n_reads=1600, n_levels=41, sticky=True, varlen=True, seed=107)
    c.append(("scale_q64S_varlen_d6_q2", v41, 6, 2, 0, (2, 0)))
    v24 = synth.block_scale(n_reads=1600, n_levels=24, sticky=True, varlen=True, seed=108, q_lo=7)
    c.append(("scale_q32S_varlen_d6_q2", v24, 6, 2, 0, (1, 0)))
    # -d0 Huffman DNA: N survives (q >= 7), symbols A,G,C,T,N are a contiguous prefix
    hn = synth.block_scale(n_levels=30, seed=109, q_lo=7, high_amb=0.02, high_codes=b"N")
    c.append(("scale_dhuff_d0_q0", hn, 0, 0, 0, (0, 1)))
    c.append(("scale_dhuff_d0_q2", hn, 0, 2, 0, (1, 1)))
    # 8-symbol DNA models at every order: N, R, W, S survive
    h8 = synth.block_scale(n_levels=12, sticky=True, seed=110, q_lo=7, high_amb=0.03, high_codes=b"NRWS")
    c.append(("scale_dna8_d3_q2", h8, 3, 2, 0, (4, 1)))
    c.append(("scale_dna8_d6_q1", h8, 6, 1, 0, (0, 1)))
    c.append(("scale_dna8_d9_q2", h8, 9, 2, 0, (4, 1)))
    # -q0 positional models: Plain (iid qualities, no tails), Truncated (iid + long '#' tails), RLE (sticky)
    c.append(("scale_plain_d0_q0", synth.block_scale(n_levels=38, seed=111, q_lo=3), 0, 0, 0, (0, 0)))
    c.append(("scale_truncated_d0_q0", synth.block_scale(n_levels=38, seed=112, q_lo=3, tail_frac=0.7, tail_max=80), 0, 0, 0, (1, 0)))
    c.append(("scale_truncated_varlen_d6_q0", synth.block_scale(n_reads=1600, n_levels=38, seed=113, q_lo=3, tail_frac=0.7, tail_max=30, varlen=True), 6, 0, 0, (1, 0)))
    c.append(("scale_rle_d6_q0", synth.block_scale(n_levels=30, sticky=True, seed=114, q_lo=3), 6, 0, 0, (2, 0)))
    # hot contexts in the large-alphabet models: a few contexts hold thousands (the partition engine's oversize partitions) or tens of
    # thousands of symbols (its fallback to the sort engine, with TSymbolCoderRC::Rescale firing inside 32- and 64-symbol rows)
    c.append(("scale_hot64_d6_q2", synth.skewed(3000, seed=115, levels=tuple(range(2, 42))), 6, 2, 0, (6, 0)))
    c.append(("scale_hot32_rescale_d6_q2", synth.skewed(5000, seed=116, levels=tuple(range(7, 31))), 6, 2, 0, (5, 0)))
    c.append(("scale_hot64_rescale_d3_q1", synth.skewed(5000, seed=117, p_major=0.99, levels=tuple(range(2, 42))), 3, 1, 0, (2, 0)))
    return c


def schemes_of(block, comp_streams):
    """(quality scheme byte, DNA scheme byte) of a compressed block: layout meta | tags | quality | dna (src/BlockCompressor.cpp:243-256),
    comp_streams in StreamsInfo order META, TAG, DNA, QUALITY"""
    meta, tag, dna, qua = [int(x) for x in comp_streams]
    return block[meta + tag], block[meta + tag + qua]


_ALL = None


def all_cases():
    """small catalogue + block-scale cases, as (name, data, dna_order, quality_order, plus_rep); generated once per process"""
    global _ALL
    if _ALL is None:
        _ALL = small_cases() + [c[:5] for c in scale_cases()]
    return _ALL
