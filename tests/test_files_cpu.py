"""Host-side file functions of the C-ABI (no device needed): PPM / PNG writers and the .rtwscene fixture format
(SURVEY.md 8f rows 1 and 2).  The PNG is decoded here by an independent reader (zlib + struct) and compared."""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest


def _read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    at, chunks = 8, []
    while at < len(data):
        n, typ = struct.unpack(">I4s", data[at:at + 8])
        body = data[at + 8:at + 8 + n]
        (crc,) = struct.unpack(">I", data[at + 8 + n:at + 12 + n])
        assert crc == zlib.crc32(typ + body) & 0xFFFFFFFF, typ
        chunks.append((typ, body))
        at += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, inter) == (8, 2, 0, 0, 0)
    raw = zlib.decompress(chunks[1][1])
    rows = np.frombuffer(raw, np.uint8).reshape(h, w * 3 + 1)
    assert not rows[:, 0].any()  # filter type 0 on every scanline
    return rows[:, 1:].reshape(h, w, 3)


@pytest.mark.parametrize("shape", [(1, 1), (54, 96), (7, 300), (225, 400)])
def test_png_and_ppm_round_trip(rtw, tmp_path, shape):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(*shape, 3), dtype=np.uint8)
    rtw.write_png(tmp_path / "a.png", img)
    assert np.array_equal(_read_png(tmp_path / "a.png"), img)
    rtw.write_ppm(tmp_path / "a.ppm", img)
    data = open(tmp_path / "a.ppm", "rb").read()
    head = f"P6\n{shape[1]} {shape[0]}\n255\n".encode()
    assert data.startswith(head) and data[len(head):] == img.tobytes()


def test_png_larger_than_one_stored_block(rtw, tmp_path):
    img = (np.arange(200 * 400 * 3, dtype=np.uint32) % 251).astype(np.uint8).reshape(200, 400, 3)  # 240 KB > 65535
    rtw.write_png(tmp_path / "b.png", img)
    assert np.array_equal(_read_png(tmp_path / "b.png"), img)


def test_rtwscene_round_trip_and_errors(rtw, tmp_path):
    rtw.reseed()
    scene = rtw.flatten_scene(rtw.scene_random_spheres())
    path = tmp_path / "random.rtwscene"
    rtw.scene_save(path, scene)
    g, m, k = rtw.scene_load(path)
    assert all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip((g, m, k), scene))
    raw = open(path, "rb").read()
    n = len(k)
    assert raw[:8] == b"RTWSCN01" and struct.unpack("<II", raw[8:16]) == (n, 0) and len(raw) == 16 + 36 * n + 4
    assert struct.unpack("<I", raw[-4:])[0] == zlib.crc32(raw[:-4]) & 0xFFFFFFFF
    # a HittableList goes through flatten_scene; an empty list is a valid file
    rtw.scene_save(tmp_path / "two.rtwscene", rtw.scene_2_spheres())
    assert len(rtw.scene_load(tmp_path / "two.rtwscene")[2]) == 2
    rtw.scene_save(tmp_path / "empty.rtwscene", [])
    assert len(rtw.scene_load(tmp_path / "empty.rtwscene")[2]) == 0
    # corruption, wrong magic, missing file, unknown material kind
    bad = bytearray(raw)
    bad[40] ^= 1
    open(tmp_path / "bad.rtwscene", "wb").write(bad)
    with pytest.raises(rtw.RtwError) as e:
        rtw.scene_load(tmp_path / "bad.rtwscene")
    assert e.value.code == rtw._lib.RTW_E_FORMAT
    open(tmp_path / "magic.rtwscene", "wb").write(b"NOTSCENE" + raw[8:])
    with pytest.raises(rtw.RtwError) as e:
        rtw.scene_load(tmp_path / "magic.rtwscene")
    assert e.value.code == rtw._lib.RTW_E_FORMAT
    with pytest.raises(rtw.RtwError) as e:
        rtw.scene_load(tmp_path / "missing.rtwscene")
    assert e.value.code == rtw._lib.RTW_E_IO
    kk = k.copy()
    kk[3] = 7
    with pytest.raises(rtw.RtwError) as e:
        rtw.scene_save(tmp_path / "kind.rtwscene", (g, m, kk))
    assert e.value.code == rtw._lib.RTW_E_UNSUPPORTED
    # too-small caller buffer: the size is still reported
    lib = rtw._lib.load()
    cnt = C.c_uint32()
    small = np.zeros((4, 4), np.float32)
    ks = np.zeros(4, np.uint32)
    fp = C.POINTER(C.c_float)
    st = lib.rtw_scene_load(str(path).encode(), small.ctypes.data_as(fp), small.ctypes.data_as(fp),
                            ks.ctypes.data_as(C.POINTER(C.c_uint32)), 4, C.byref(cnt))
    assert st == rtw._lib.RTW_E_INVALID_ARG and cnt.value == n


def test_oracle_reads_the_same_scene_file(rtw, oracle, tmp_path):
    # the fixture format is shared: a scene saved by the library renders identically through the oracle after loading
    scene = rtw.flatten_scene(rtw.scene_4_spheres())
    rtw.scene_save(tmp_path / "four.rtwscene", scene)
    loaded = rtw.scene_load(tmp_path / "four.rtwscene")
    cam = rtw.t_default_cam()
    a, _, _ = oracle.render(*scene, cam.as_array(), 32, 2, max_depth=4, seed=1, n_threads=1)
    b, _, _ = oracle.render(*loaded, cam.as_array(), 32, 2, max_depth=4, seed=1, n_threads=1)
    assert np.array_equal(a, b)
