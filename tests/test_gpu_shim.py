"""The drop-in itself: host/BlockCompressorGpu.h (the C++ class a DSRC maintainer puts in place of comp::BlockCompressor,
INTEGRATION.md binding 1), compiled against the reference's own headers by oracle/Makefile (target `shim`), against the
UNMODIFIED reference's BlockCompressor driven through identical calls -- first ("cold") and second ("warm") Store of one
instance (SURVEY 8-Q1 lives in the shim's tagCapacity), StreamsInfo, and Read in both directions (each side decodes the
other's block)."""
import pytest

import cases
import refbind

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refbind.ref_available() and refbind.shim_available()), reason="oracle/_ref harnesses not built")]

NAMES = ["ill_binned_d6_q2", "ill_full_d6_q2", "ion454_d9_q2", "ill_binned_d0_q0", "ill_plusrep_d6_q2", "mixed_titles_d6_q2",
         "q1_valuevar_d6_q2", "ill_mid12_d3_q1", "tiny2_d6_q2"]
CASES = [c for c in cases.small_cases() if c[0] in NAMES]


@pytest.mark.parametrize("name,data,d,q,pr", CASES, ids=[c[0] for c in CASES])
def test_shim_is_a_drop_in_for_block_compressor(name, data, d, q, pr):
    chunk = data[:-1]
    ref = refbind.Ref(33, pr, d, q)
    gpu = refbind.Shim(33, pr, d, q, max_block=len(chunk) + 64)
    for call in ("cold", "warm"):
        exp, eraw, ecmp = ref.store(chunk)
        got, graw, gcmp = gpu.store(chunk)
        assert got == exp, "%s Store differs" % call
        assert (graw, gcmp) == (eraw, ecmp)
    assert gpu.read(exp) == data          # the GPU decodes the reference's block
    assert ref.read(got) == data          # the reference decodes the GPU's block


def test_shim_crc_blocks():
    name, data, d, q, pr = [c for c in cases.small_cases() if c[0] == "ill_binned_d6_q2"][0]
    chunk = data[:-1]
    ref = refbind.Ref(33, pr, d, q, crc=True)
    gpu = refbind.Shim(33, pr, d, q, crc=True, max_block=len(chunk) + 64)
    exp, _, _ = ref.store(chunk)
    got, _, _ = gpu.store(chunk)
    assert got == exp
    assert gpu.read(exp) == data
