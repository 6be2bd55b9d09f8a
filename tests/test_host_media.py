"""The reference's own input media read without OpenCV (host/lm_media.hpp): uncompressed AVI video (what
cv::VideoCapture + extractChannel(0) deliver, LocoMouse_class.cpp:367-400, 1282-1293) and the PNG background
(imread(..., CV_LOAD_IMAGE_GRAYSCALE), 402-417), each against what the REAL OpenCV reads from the same file."""
import ctypes as C
import struct

import numpy as np
import pytest

from test_match2nd import _host

cv2 = pytest.importorskip("cv2")


def _read(fn_name, path):
    H = _host()
    fn = getattr(H, fn_name)
    fn.restype = C.c_int
    dims = np.zeros(3, np.int32)
    msg = C.create_string_buffer(512)
    rc = fn(str(path).encode(), C.c_void_p(dims.ctypes.data), None, msg, 512)
    if rc != 0:
        raise RuntimeError(msg.value.decode())
    out = np.zeros(tuple(int(v) for v in dims), np.uint8)
    assert fn(str(path).encode(), C.c_void_p(dims.ctypes.data), C.c_void_p(out.ctypes.data), msg, 512) == 0
    return out


def test_grey_avi_written_by_opencv(tmp_path):
    rng = np.random.Generator(np.random.PCG64(1))
    frames = rng.integers(0, 256, (7, 50, 68), dtype=np.uint8)
    w = cv2.VideoWriter(str(tmp_path / "v.avi"), 0, 30.0, (68, 50), False)   # fourcc 0, isColor False -> 'Y800'
    assert w.isOpened()
    for f in frames:
        w.write(f)
    w.release()
    cap = cv2.VideoCapture(str(tmp_path / "v.avi"))
    want = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        want.append(cv2.extractChannel(f, 0))
    want = np.stack(want)
    got = _read("lmh_read_avi", tmp_path / "v.avi")
    assert np.array_equal(want, frames) and np.array_equal(got, want)


def _riff(tag, payload):
    return tag + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")


def _dib_avi(frames_bgr_or_idx, bits, palette=None, top_down=False):
    n, h, w = frames_bgr_or_idx.shape[:3]
    stride = (w * bits // 8 + 3) & ~3
    bih = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bits, 0, stride * h, 0, 0, len(palette) if palette is not None else 0, 0)
    if palette is not None:
        bih += b"".join(struct.pack("<BBBB", b, g, r, 0) for (b, g, r) in palette)
    strh = b"vids" + b"DIB " + b"\0" * 48
    hdrl = b"hdrl" + _riff(b"avih", b"\0" * 56) + _riff(b"LIST", b"strl" + _riff(b"strh", strh) + _riff(b"strf", bih))
    movi = b"movi"
    for f in frames_bgr_or_idx:
        rows = f if top_down else f[::-1]
        data = b"".join(np.ascontiguousarray(r).tobytes() + b"\0" * (stride - w * bits // 8) for r in rows)
        movi += _riff(b"00db", data)
    body = b"AVI " + _riff(b"LIST", hdrl) + _riff(b"JUNK", b"\0" * 11) + _riff(b"LIST", movi)
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_bi_rgb_avi_variants(tmp_path):
    """Hand-built DIB AVIs (24-bit bottom-up, 32-bit top-down, 8-bit palette, odd width with row padding): channel 0 (blue),
    as extractChannel(F, F, 0) takes it from the BGR frame VideoCapture delivers."""
    rng = np.random.Generator(np.random.PCG64(2))
    bgr = rng.integers(0, 256, (3, 9, 13, 3), dtype=np.uint8)
    (tmp_path / "a.avi").write_bytes(_dib_avi(bgr, 24))
    got = _read("lmh_read_avi", tmp_path / "a.avi")
    assert np.array_equal(got, bgr[..., 0])
    bgra = rng.integers(0, 256, (2, 6, 10, 4), dtype=np.uint8)
    (tmp_path / "b.avi").write_bytes(_dib_avi(bgra, 32, top_down=True))
    assert np.array_equal(_read("lmh_read_avi", tmp_path / "b.avi"), bgra[..., 0])
    idx = rng.integers(0, 256, (2, 5, 7), dtype=np.uint8)
    pal = [(int(rng.integers(0, 256)), 0, 0) for _ in range(256)]
    (tmp_path / "c.avi").write_bytes(_dib_avi(idx, 8, palette=pal))
    assert np.array_equal(_read("lmh_read_avi", tmp_path / "c.avi"), np.array([p[0] for p in pal], np.uint8)[idx])


def test_compressed_video_is_refused_with_a_reason(tmp_path):
    w = cv2.VideoWriter(str(tmp_path / "m.avi"), cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (64, 48), True)
    if not w.isOpened():
        pytest.skip("no MJPG encoder in this OpenCV build")
    w.write(np.zeros((48, 64, 3), np.uint8))
    w.release()
    with pytest.raises(RuntimeError, match="outside this library"):
        _read("lmh_read_avi", tmp_path / "m.avi")


@pytest.mark.parametrize("kind", ["gray", "bgr", "bgra", "gray_equal_channels"])
def test_png_background_equals_imread_grayscale(tmp_path, kind):
    rng = np.random.Generator(np.random.PCG64(3))
    h, w = 37, 53
    if kind == "gray":
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img[5:20, 7:30] = np.arange(23, dtype=np.uint8)[None, :] * 3   # smooth ramps exercise the Sub / Up / Paeth row filters
    elif kind == "bgr":
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == "bgra":
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    else:
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img = np.stack([g, g, g], -1)
    p = tmp_path / "bkg.png"
    assert cv2.imwrite(str(p), img)
    want = cv2.imread(str(p), cv2.IMREAD_GRAYSCALE)
    got = _read("lmh_read_png", p)[0]
    assert got.shape == want.shape
    assert np.array_equal(got, want)
