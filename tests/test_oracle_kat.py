"""CPU tests that pin the oracle: the README's worked examples (README.md:45,53-66,72,80,86-96), the libstdc++ random
number known answers, the A-buffer byte count quoted in the reference's screenshot, and table identities.
The reference ships no golden images or vectors (test.py only checks the exit code), so these are all there is."""
import ctypes as C
import struct

import numpy as np
import pytest


def fbits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


C1, C2, C3, C4 = (0.9, 0.1, 0.1, 0.5), (0.1, 0.8, 0.2, 0.4), (0.2, 0.3, 0.9, 0.6), (0.7, 0.7, 0.1, 0.3)
README_FRAGS = [(C4, 0.4), (C2, 0.2), (C1, 0.1), (C3, 0.3)]


def pack(O, c):
    return O.lib().oracle_pack_color((C.c_float * 4)(*c))


def premult(c):
    c = np.float32(c)
    return np.array([c[0] * c[3], c[1] * c[3], c[2] * c[3], c[3]], np.float32)


def unpack_premult(O, p):
    d = [O.lib().oracle_srgb_decode8((p >> s) & 255) for s in (0, 8, 16)]
    a = np.float32(p >> 24) / np.float32(255.0)
    return np.array([np.float32(d[0]) * a, np.float32(d[1]) * a, np.float32(d[2]) * a, a], np.float32)


def make(O, alg, L=2, W=4, H=4, N=10, tail=1):
    o = O.Oracle(O.make_config(algorithm=alg, oitLayers=L, linkedListAllocatedPerElement=N, tailBlend=tail, width=W, height=H))
    o.set_scene_data(O.camera(W, H))
    o.begin_frame()
    return o


def test_readme_simple(oracle_mod):
    O = oracle_mod
    o = make(O, O.OIT_SIMPLE)
    outs = [o.debug_invoke(1, 2, c, z) for c, z in README_FRAGS]
    vs = 16
    ab = o.abuffer.reshape(-1, 2)
    pos = 2 * 4 + 1
    assert tuple(ab[pos]) == (pack(O, C4), fbits(0.4))          # first two fragments are stored, unsorted
    assert tuple(ab[pos + vs]) == (pack(O, C2), fbits(0.2))
    assert not outs[0].any() and not outs[1].any()
    assert np.array_equal(outs[2], premult(C1)) and np.array_equal(outs[3], premult(C3))  # c1 then c3 tail blended
    assert o.aux(0)[pos] == 4                                      # the counter keeps counting past OIT_LAYERS
    before = o.color_samples[2, 1, 0]
    o.composite()
    # composite sorts to (c2, 0.2), (c4, 0.4) and blends front to back
    exp = unpack_premult(O, pack(O, C2)).astype(np.float64)
    b = unpack_premult(O, pack(O, C4)).astype(np.float64)
    exp = exp + (1 - exp[3]) * b
    got = O.lib().oracle_rop_blend(int(before), (C.c_float * 4)(*exp.astype(np.float32)))
    assert abs(int(o.color_samples[2, 1, 0] & 255) - int(got & 255)) <= 1
    assert o.stats["fragmentsStored"] == 2 and o.stats["fragmentsTail"] == 2


def test_readme_linked_list_table(oracle_mod):
    O = oracle_mod
    o = make(O, O.OIT_LINKEDLIST, W=2, H=2, N=1)  # pool of 4 nodes including the terminator
    p1, p2 = (0, 0), (1, 0)
    outs = [o.debug_invoke(*p1, C4, 0.4), o.debug_invoke(*p2, C1, 0.1), o.debug_invoke(*p1, C2, 0.2), o.debug_invoke(*p1, C3, 0.3)]
    ab = o.abuffer.reshape(-1, 4)
    assert tuple(ab[1]) == (pack(O, C4), fbits(0.4), 0, 0)
    assert tuple(ab[2]) == (pack(O, C1), fbits(0.1), 0, 0)
    assert tuple(ab[3]) == (pack(O, C2), fbits(0.2), 0, 1)
    assert o.aux(0)[0] == 3 and o.aux(0)[1] == 2                    # list heads
    assert np.array_equal(outs[3], premult(C3))                    # (c3, 0.3) ran out of memory: newOffset(4) >= 4
    assert o.aux(3)[0] == 4


def test_readme_loop32(oracle_mod):
    O = oracle_mod
    o = make(O, O.OIT_LOOP)
    for c, z in README_FRAGS:
        o.debug_invoke(0, 0, c, z, pass_=0)
    ab, vs = o.abuffer, 16
    assert (ab[0], ab[vs]) == (fbits(0.1), fbits(0.2))             # frontmost sorted depths
    outs = [o.debug_invoke(0, 0, c, z, pass_=1) for c, z in README_FRAGS]
    assert (ab[2 * vs], ab[3 * vs]) == (pack(O, C1), pack(O, C2))  # colours matched to depths
    assert np.array_equal(outs[0], premult(C4)) and np.array_equal(outs[3], premult(C3))  # tail: c4 and c3
    assert not outs[1].any() and not outs[2].any()


def test_readme_loop64(oracle_mod):
    O = oracle_mod
    o = make(O, O.OIT_LOOP64)
    outs = [o.debug_invoke(0, 0, c, z) for c, z in README_FRAGS]
    ab = o.abuffer.reshape(-1, 2)
    assert tuple(ab[0]) == (pack(O, C1), fbits(0.1)) and tuple(ab[16]) == (pack(O, C2), fbits(0.2))
    # the evicted (c4) and the rejected (c3) fragment are tail blended with their 8-bit quantised colours
    assert np.array_equal(outs[2], unpack_premult(O, pack(O, C4)))
    assert np.array_equal(outs[3], unpack_premult(O, pack(O, C3)))


@pytest.mark.parametrize("alg", ["spinlock", "interlock"])
def test_readme_spinlock_trace(oracle_mod, alg):
    O = oracle_mod
    o = make(O, O.OIT_SPINLOCK if alg == "spinlock" else O.OIT_INTERLOCK)
    # critical-section order of the README trace: c3, c2, (c4 rejected), c1 evicts c3
    outs = [o.debug_invoke(0, 0, c, z) for c, z in [(C3, 0.3), (C2, 0.2), (C4, 0.4), (C1, 0.1)]]
    ab = o.abuffer.reshape(-1, 2)
    assert tuple(ab[0]) == (pack(O, C1), fbits(0.1)) and tuple(ab[16]) == (pack(O, C2), fbits(0.2))
    assert np.array_equal(outs[2], premult(C4))                    # rejected: tail blends its own float colour
    assert np.array_equal(outs[3], unpack_premult(O, pack(O, C3)))  # evicted: the decoded 8-bit colour
    assert o.aux(2)[0] == fbits(0.3)                               # imgDepth holds the evicted depth
    assert o.aux(0)[0] == 4


def test_rng_known_answers(oracle_mod):
    O = oracle_mod
    s = C.c_uint64(3625)
    got = [O.lib().oracle_rand_canonical(C.byref(s)) for _ in range(8)]
    want = [0.028370589, 0.824482024, 0.0692929551, 0.606728196, 0.281235605, 0.727050424, 0.536733806, 0.884598672]
    assert np.allclose(got, want, rtol=0, atol=5e-10)
    verts, idx, ipo = O.generate_scene(O.make_config(numObjects=2, subdiv=16))
    # sphere 0: x = 3rd draw, y = 2nd, z = 1st (g++ evaluates the constructor arguments right to left)
    centre = (np.float32([0.0692929551, 0.824482024, 0.028370589]) - np.float32(0.5)) * np.float32(8)
    poles = (verts[0, :3] + verts[560, :3]) / 2  # first / last vertex are the +z / -z poles
    assert np.allclose(poles, centre, atol=1e-5)
    assert verts.shape == (2 * 561, 10) and ipo == 960 * 3 and idx.size == 2 * 960 * 3  # main.cpp:346, SURVEY 2.3
    assert idx.max() == 2 * 561 - 2  # the duplicated seam vertex of the last ring is never referenced


def test_abuffer_size_matches_reference_screenshot(oracle_mod):
    O = oracle_mod
    o = O.Oracle(O.make_config(algorithm=O.OIT_INTERLOCK, oitLayers=16, aaType=O.AA_MSAA_4X, width=1920, height=1017))
    assert o.abuffer.nbytes == 499875840  # doc/vk_order_independent_transparency.png, oit.cpp:155-156
    o.close()


def test_srgb_tables_round_trip(oracle_mod):
    L = oracle_mod.lib()
    for v in range(256):
        assert L.oracle_srgb_encode8(L.oracle_srgb_decode8(v)) == v
    assert L.oracle_srgb_encode8(0.2) == 124 and L.oracle_srgb_encode8(-1.0) == 0 and L.oracle_srgb_encode8(7.0) == 255
    # a zero-colour blend is the identity on every destination byte pattern component
    zero = (C.c_float * 4)(0, 0, 0, 0)
    for v in range(256):
        d = v | (v << 8) | (v << 16) | (v << 24)
        assert L.oracle_rop_blend(d, zero) == d


def test_half_conversions(oracle_mod):
    L = oracle_mod.lib()
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32) * 100, np.float32([0, 1, 65504, 65520, 1e-8, 6e-8, 3e-5, 1e6]),
                           rng.random(2000).astype(np.float32) * 1e-4])
    for v in vals:
        want = np.float16(v).view(np.uint16)
        assert L.oracle_float_to_half(float(v)) == int(want), v
    for h in range(0, 0x7C00, 37):
        assert L.oracle_half_to_float(h) == float(np.uint16(h).view(np.float16))
