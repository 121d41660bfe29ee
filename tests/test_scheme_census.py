"""Scheme census of the parity catalogue: every (order, scheme) pair the reference can select on the lossless path must be hit by at
least one case (small catalogue + block-scale cases), and the block-scale cases must land on the scheme they were built for. The
scheme bytes are read from the block the oracle writes (pinned to the reference by test_oracle_vs_ref.py); the GPU parity tests
run the same catalogue, so this census is also the GPU's.
    quality: -q0 {0 Plain, 1 Truncated, 2 RLE} (src/QualityModelerProxy.h:113-122); -q1 {0..3}; -q2 {0..7} (:231-282)
    DNA:     -d0 {0 2-bit, 1 Huffman} (src/DnaModelerProxy.h:102-111); -d1/2/3 {0 4-symbol, 1 8-symbol} (:160-170)"""
import cases
import refbind

WANT_Q = {(0, 0), (0, 1), (0, 2)} | {(1, k) for k in range(4)} | {(2, k) for k in range(8)}
WANT_D = {(o, k) for o in (0, 3, 6, 9) for k in (0, 1)}


def _census(catalogue):
    hit_q, hit_d = {}, {}
    for name, data, d, q, pr in catalogue:
        blk, _, cmp_ = refbind.Oracle(33, pr, d, q).store(data[:-1])
        qs, ds = cases.schemes_of(blk, cmp_)
        hit_q.setdefault((q, qs), []).append((name, len(data)))
        hit_d.setdefault((d, ds), []).append((name, len(data)))
    return hit_q, hit_d


def test_every_scheme_is_hit_and_at_block_scale():
    hit_q, hit_d = _census(cases.all_cases())
    assert WANT_Q <= set(hit_q), sorted(WANT_Q - set(hit_q))
    assert WANT_D <= set(hit_d), sorted(WANT_D - set(hit_d))
    # ... and by at least one input of >= 256 KiB, so the engines cross many tiles on it
    for k in sorted(WANT_Q):
        assert any(n >= (1 << 18) for _, n in hit_q[k]), ("quality", k)
    for k in sorted(WANT_D):
        assert any(n >= (1 << 18) for _, n in hit_d[k]), ("dna", k)


def test_scale_cases_select_the_scheme_they_were_built_for():
    for name, data, d, q, pr, want in cases.scale_cases():
        blk, _, cmp_ = refbind.Oracle(33, pr, d, q).store(data[:-1])
        assert cases.schemes_of(blk, cmp_) == want, name
