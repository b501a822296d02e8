"""The host-side packer of the packed upload (csrc/host_pack.cpp, through dsp_debug_pack_chunk; no GPU needed): every form against
a numpy restatement of the formats in csrc/host_pack.hpp, the AVX2 packer against the portable one, and the thresholds between the
forms.  What is packed is Chunk::blocks, 4,096 u16 (/root/reference/src/content/world/chunks/chunk.rs:17)."""
import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def unpack(desc, packed):
    form = desc & 7
    if form == 0:
        return np.zeros(4096, dtype=np.uint16)
    if form == 1:
        return np.full(4096, desc >> 16, dtype=np.uint16)
    if form == 5:
        return packed.view(np.uint16).copy()
    bits = {2: 2, 3: 4, 4: 8}[form]
    i = np.arange(4096)
    per = 8 // bits
    return ((packed[i // per].astype(np.uint16) >> (bits * (i % per))) & ((1 << bits) - 1)).astype(np.uint16)


def chunks():
    rng = np.random.default_rng(5)
    out = {"air": np.zeros(4096, dtype=np.uint16)}
    for uid in (1, 3, 255, 256, 65535):
        out[f"uniform {uid}"] = np.full(4096, uid, dtype=np.uint16)
    for top, form in ((3, 2), (4, 3), (15, 3), (16, 4), (255, 4), (256, 5), (65535, 5)):
        c = rng.integers(0, top + 1, 4096).astype(np.uint16)
        c[rng.integers(0, 4096)] = top  # the largest id is present: the form is decided by it
        c[0] = 0 if c[1] else 1         # never uniform
        out[f"random up to {top}"] = (c, form)
    one = np.zeros(4096, dtype=np.uint16)
    one[4095] = 2
    out["single voxel, last position"] = (one, 2)
    last_differs = np.full(4096, 7, dtype=np.uint16)
    last_differs[4095] = 6
    out["uniform but for the last voxel"] = (last_differs, 3)
    sparse_big = np.zeros(4096, dtype=np.uint16)
    sparse_big[2049] = 300
    out["one id beyond 8 bits"] = (sparse_big, 5)
    return out


@pytest.mark.parametrize("name", list(chunks()))
def test_packed_forms_are_lossless_and_minimal(name):
    c = chunks()[name]
    want_form = None
    if isinstance(c, tuple):
        c, want_form = c
    d_fast, p_fast = capi.debug_pack_chunk(c)
    d_port, p_port = capi.debug_pack_chunk(c, portable=True)
    assert d_fast == d_port and np.array_equal(p_fast, p_port)
    assert np.array_equal(unpack(d_fast, p_fast), c)
    if name == "air":
        assert d_fast & 7 == 0
    elif name.startswith("uniform ") and want_form is None:
        assert d_fast & 7 == 1 and d_fast >> 16 == int(c[0])
    else:
        assert d_fast & 7 == want_form
    assert len(p_fast) == (0, 0, 1024, 2048, 4096, 8192)[d_fast & 7]


def test_packer_on_many_random_chunks():
    rng = np.random.default_rng(9)
    for k in range(300):
        top = int(rng.choice([1, 2, 3, 5, 15, 40, 255, 1000, 65535]))
        density = rng.random()
        c = ((rng.random(4096) < density) * rng.integers(1, top + 1, 4096)).astype(np.uint16)
        d, p = capi.debug_pack_chunk(c)
        assert np.array_equal(unpack(d, p), c), (k, top)
        assert (d, p.tobytes()) == (lambda r: (r[0], r[1].tobytes()))(capi.debug_pack_chunk(c, portable=True))
