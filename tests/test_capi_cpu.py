"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the
header declares, fails loudly without a GPU, and the host bookkeeping (GalleryDict,
shard partitioning) behaves like the reference's dict."""
import ctypes
import os
import re

import numpy as np
import pytest

import frp_b200
from frp_b200 import _capi
from frp_b200.gallery import GalleryDict
from frp_b200.shard import record_bytes, shard_bounds, shard_streams, unpack_records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "frp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 28
    lib = ctypes.CDLL(_capi._LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libfrp_b200.so does not export %s" % n
    assert set(names) == set(_capi.SIGNATURES.keys())


def test_abi_version_and_record_size():
    assert _capi.lib.frp_abi_version() == 2
    for B, k in ((1, 1), (3, 10), (256, 10), (1024, 64)):
        assert _capi.lib.frp_shard_record_bytes(B, k) == record_bytes(B, k)


def test_invalid_arguments_are_rejected_without_a_device():
    h = ctypes.c_void_p()
    assert _capi.lib.frp_gallery_create(0, 0, 0, 10, ctypes.byref(h)) == _capi.FRP_E_INVALID
    assert _capi.lib.frp_gallery_create(0, 128, 7, 10, ctypes.byref(h)) == _capi.FRP_E_INVALID
    assert "invalid" in _capi.last_error()
    assert _capi.lib.frp_gallery_count(None, None, None) == _capi.FRP_E_INVALID
    assert _capi.lib.frp_match(None, None, 1, 0.6, 5, 0, None, None, None, None) == _capi.FRP_E_INVALID


@pytest.mark.skipif(frp_b200.device_count() > 0, reason="a GPU is visible")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(frp_b200.FrpError) as e:
        frp_b200.Gallery(128)
    assert e.value.code == _capi.FRP_E_NODEVICE
    fr = np.zeros((1, 32, 32, 3), np.uint8)
    with pytest.raises(frp_b200.FrpError):
        frp_b200.align.norm_crop_batch(fr, [0], lmk=np.zeros((1, 5, 2), np.float32))
    d = GalleryDict()
    with pytest.raises(frp_b200.FrpError):
        d["a"] = [0.0] * 128


class _HostGallery:
    """Host stand-in used ONLY to test GalleryDict's bookkeeping on CPU."""

    def __init__(self, dim, metric, device):
        self.dim, self.metric, self.device = dim, metric, device
        self.data, self.alive = [], []

    def append(self, rows):
        first = len(self.data)
        for r in np.atleast_2d(rows):
            self.data.append(np.array(r, np.float64))
            self.alive.append(True)
        return first

    def update(self, row, vec):
        assert self.alive[row]
        self.data[row] = np.array(vec, np.float64)

    def remove(self, row):
        assert self.alive[row]
        self.alive[row] = False

    def gather(self, idx):
        g = _HostGallery(self.dim, self.metric, self.device)
        g.append(np.stack([self.data[int(i)] for i in idx]))
        return g

    def get_row(self, row):
        return self.data[row].astype(np.float32)

    def close(self):
        pass


def test_gallerydict_mirrors_python_dict_order():
    rng = np.random.default_rng(0)
    d = GalleryDict(gallery_factory=_HostGallery, compact_ratio=0.25)
    ref = {}
    names = ["n%d" % i for i in range(40)]
    for step in range(4000):
        n = names[rng.integers(len(names))]
        op = rng.random()
        if op < 0.55:
            v = rng.normal(size=8).tolist()
            d[n] = v
            ref[n] = v
        elif op < 0.9:
            if n in ref:
                del d[n]
                del ref[n]
            else:
                with pytest.raises(KeyError):
                    del d[n]
        else:
            assert d.pop(n, None) == ref.pop(n, None)
        if step % 97 == 0:
            assert d.keys() == list(ref.keys())
            assert len(d) == len(ref)
            # row order == dict order, every live row holds its vector
            rows = [d.row_of(k) for k in d.keys()]
            assert rows == sorted(rows)
            for k in ref:
                assert d[k] == ref[k]
                assert np.allclose(d.gallery.data[d.row_of(k)], ref[k])
                assert d.name_of(d.row_of(k)) == k
    d.compact()
    assert d.keys() == list(ref.keys())
    assert [d.row_of(k) for k in d.keys()] == list(range(len(ref)))
    d.clear()
    assert len(d) == 0 and d.keys() == [] and not d


def test_gallerydict_rejects_dim_mismatch_and_bulk_load():
    d = GalleryDict(gallery_factory=_HostGallery)
    d.bulk_load(["a", "b", "c"], np.eye(3, 4))
    assert d.keys() == ["a", "b", "c"] and d["b"] == [0.0, 1.0, 0.0, 0.0]
    with pytest.raises(ValueError):
        d["x"] = [1.0, 2.0]
    with pytest.raises(ValueError):
        d.bulk_load(["a"], np.zeros((1, 4)))


def test_cooldown_filter_matches_reference_semantics():
    """tracking_service.py:122-134,179: a (person, camera) pair is recorded again only
    once the cooldown has elapsed since the last RECORDED detection."""
    from frp_b200.camera_scan import CooldownFilter
    f = CooldownFilter(10)
    det = lambda t, c: {"camera_id": c, "target": t, "distance": 0.3, "confidence": "high"}
    r0 = f.filter({1: [det("a", 1), det("b", 1)], 2: [det("a", 2)]}, now=100.0)
    assert [d["target"] for d in r0[1]] == ["a", "b"] and len(r0[2]) == 1
    r1 = f.filter({1: [det("a", 1)], 2: [det("c", 2)]}, now=105.0)
    assert r1[1] == [] and [d["target"] for d in r1[2]] == ["c"]
    r2 = f.filter({1: [det("a", 1)]}, now=109.9)          # still measured from t=100, not 105
    assert r2[1] == []
    r3 = f.filter({1: [det("a", 1)]}, now=110.0)          # >= cooldown: recorded again
    assert len(r3[1]) == 1


def test_shard_bounds_partition():
    for total in (0, 1, 7, 100, 10_000_019):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_streams_partition():
    cams = ["cam%02d" % i for i in (7, 3, 11, 0, 5, 9, 1)]
    for world in (1, 2, 4, 8):
        parts = [shard_streams(cams, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(cams)              # every stream exactly once
        assert max(map(len, parts)) - min(map(len, parts)) <= 1    # balanced
        assert parts == [shard_streams(list(reversed(cams)), world, r) for r in range(world)]   # order-free
    assert shard_streams([3, 1, 2], 2, 0) == [1, 3] and shard_streams([3, 1, 2], 2, 1) == [2]


def test_record_pack_roundtrip():
    B, k, R = 5, 3, 2
    nb = record_bytes(B, k)
    buf = np.zeros(R * nb, np.uint8)
    for r in range(R):
        rec = buf[r * nb:(r + 1) * nb]
        rec[:B * k * 8].view(np.int64)[:] = np.arange(B * k) + 100 * r
        rec[B * k * 8:B * k * 16].view(np.float64)[:] = np.arange(B * k) * 0.5 + r
        rec[B * k * 16:B * k * 16 + B * 4].view(np.int32)[:] = r + 1
    idx, sc, cnt, st = unpack_records(buf, R, B, k)
    assert idx.shape == (R, B, k) and idx[1, 0, 0] == 100 and sc[1, 0, 1] == 1.5 and cnt[1, 2] == 2


def test_built_library_contains_the_blackwell_instructions():
    """The shipped .so is sm_100a code with the instructions the design names (no library or
    generic fallback): tcgen05 MMAs incl. the cta_group::2 form, TMEM loads, tiled and bulk TMA,
    multicast commits, and the integer dot product of the alignment blend."""
    import shutil
    import subprocess
    from frp_b200 import _capi
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _capi._LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper()
    for op in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG.2D", "UTMALDG.2D.2CTA", "UTMALDG.2D.MULTICAST",
               "UTMALDG.3D", "UBLKCP", "UTCBAR.2CTA.MULTICAST", "IDP.2A"):
        assert op in sass, op


def test_ptr_returns_the_array_address_on_every_route():
    """_capi.ptr: the buffer-protocol fast route and the ndarray.ctypes fallback (read-only and
    empty arrays) give the same address; views carry their offset; strided arrays are refused."""
    import ctypes as C
    from frp_b200._capi import ptr
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    assert ptr(a) == a.ctypes.data
    assert C.cast(ptr(a), C.POINTER(C.c_float))[5] == 5.0
    ro = np.frombuffer(b"\x00" * 16, np.float32)
    assert not ro.flags.writeable and ptr(ro) == ro.ctypes.data
    empty = np.empty((0, 4), np.float32)
    assert ptr(empty) == empty.ctypes.data
    buf = np.empty(96, np.uint8)
    idx = np.ndarray((2, 5), np.int64, buf, 0)
    cnt = np.ndarray((2,), np.int32, buf, 80)
    assert ptr(idx) == ptr(buf) and ptr(cnt) == ptr(buf) + 80
    with pytest.raises(ValueError):
        ptr(a[:, ::2])
    assert ptr(None) is None and ptr(1234) == 1234
    with pytest.raises(TypeError):
        ptr("nope")
