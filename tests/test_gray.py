"""grayscale of camera frames (CompVImage::convertGrayscale, base/image/compv_image.cxx:687-692 -> compv_image_conv_to_grayscale.cxx:35-93).
A numpy restatement of the reference's integer arithmetic is pinned on the compiled reference (CPU); the CUDA kernel and the pipeline entry point that takes camera
formats are compared with it bit for bit (GPU)."""
import numpy as np
import pytest

import oracle
from frames import frame_g

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")

FMT = dict(RGB24=13, BGR24=14, RGBA32=15, BGRA32=16, ABGR32=17, ARGB32=18, RGB565LE=19, RGB565BE=20, BGR565LE=21, BGR565BE=22, Y=25, NV12=26, NV21=27, YUV420P=28, YVU420P=29, YUV422P=30,
           YUYV422=31, UYVY422=32, YUV444P=33)
BPP = dict(RGB24=3, BGR24=3, RGBA32=4, BGRA32=4, ARGB32=4, RGB565LE=2, RGB565BE=2, BGR565LE=2, BGR565BE=2, YUYV422=2, UYVY422=2)
PLANAR_EXTRA = dict(Y=0.0, NV12=0.5, NV21=0.5, YUV420P=0.5, YVU420P=0.5, YUV422P=1.0, YUV444P=2.0)  # chroma bytes per luma byte


def make_frame(name, w, h, stride, seed):
    """A whole frame buffer in the layout CompVImage uses: `stride` samples per row; planar formats: Y plane then chroma (its content does not matter for gray)."""
    rng = np.random.default_rng(seed)
    if name in BPP:
        buf = rng.integers(0, 256, (h, stride * BPP[name]), dtype=np.uint8)
        return buf.reshape(-1)
    y = np.zeros((h, stride), np.uint8)
    y[:, :w] = frame_g(w, h, seed)
    y[:, w:] = 77
    chroma = rng.integers(0, 256, int(h * stride * PLANAR_EXTRA[name]) + 64, dtype=np.uint8)
    return np.concatenate([y.reshape(-1), chroma])


def gray_restated(name, buf, w, h, stride):
    """compv_image_conv_rgbfamily.cxx:93-120, 243-270, 400-425; compv_image_conv_to_grayscale.cxx:260-280; integer arithmetic as written there."""
    if name not in BPP:
        return buf[:h * stride].reshape(h, stride)[:, :w].copy()
    bpp = BPP[name]
    px = buf[:h * stride * bpp].reshape(h, stride, bpp)[:, :w].astype(np.int32)
    if name in ("YUYV422", "UYVY422"):
        return px[:, :, 0 if name == "YUYV422" else 1].astype(np.uint8)
    if bpp == 2:
        k = (px[:, :, 0] | (px[:, :, 1] << 8)) if name.endswith("LE") else ((px[:, :, 0] << 8) | px[:, :, 1])
        r = (k & 0xF800) >> 8; r |= r >> 5
        g = (k & 0x07E0) >> 3; g |= g >> 6
        b = (k & 0x001F) << 3; b |= b >> 5
        c = (33, 65, 13) if name.startswith("RGB") else (13, 65, 33)
        return np.minimum(((c[0] * r + c[1] * g + c[2] * b) >> 7) + 16, 255).astype(np.uint8)
    off = 1 if name == "ARGB32" else 0
    c = (13, 65, 33) if name.startswith("BGR") else (33, 65, 13)
    return np.minimum(((c[0] * px[:, :, off] + c[1] * px[:, :, off + 1] + c[2] * px[:, :, off + 2]) >> 7) + 16, 255).astype(np.uint8)


CONVERTIBLE = [n for n in FMT if n != "ABGR32"]


@needs_ref
@pytest.mark.parametrize("name", CONVERTIBLE)
def test_gray_restatement_equals_the_reference(name):
    for (w, h, stride) in [(64, 48, 64), (130, 41, 160), (640, 480, 640)]:
        buf = make_frame(name, w, h, stride, 3)
        np.testing.assert_array_equal(oracle.to_grayscale_ref(FMT[name], buf, w, h, stride), gray_restated(name, buf, w, h, stride))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CONVERTIBLE)
def test_cuda_gray(cvb, name):
    for (w, h, stride) in [(64, 48, 64), (130, 41, 160), (1920, 1080, 1920), (1001, 77, 1024)]:
        buf = make_frame(name, w, h, stride, 5)
        got = cvb.image_to_grayscale(FMT[name], buf, w, h, stride)
        np.testing.assert_array_equal(got[:, :w], gray_restated(name, buf, w, h, stride))


@pytest.mark.gpu
def test_cuda_gray_rejects_what_the_reference_rejects(cvb):
    from compv_b200 import _ffi
    with pytest.raises(_ffi.CvbError) as e:
        cvb.image_to_grayscale(FMT["ABGR32"], np.zeros(64 * 48 * 4, np.uint8), 64, 48, 64)
    assert e.value.code == _ffi.E_NOT_IMPLEMENTED          # compv_image_conv_to_grayscale.cxx:88-91


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["RGB24", "BGRA32", "NV12", "UYVY422", "RGB565LE"])
def test_cuda_canny_kht_pipeline_on_camera_formats(cvb, name):
    """cvb200_canny_kht_process_batch_fmt: raw camera frames in, lines out; must equal gray conversion (restated) -> Gaussian -> Canny -> KHT of the oracle."""
    from compv_b200 import _ffi
    w, h, stride, batch = 320, 200, 320, 7
    bufs = [make_frame(name, w, h, stride, 40 + k) for k in range(batch)]
    if name in BPP:
        # random colour noise has no structure: paint the structured gray frame into the channels so that there are lines to find
        for k, b in enumerate(bufs):
            g = frame_g(w, h, 40 + k)
            v = b[:h * stride * BPP[name]].reshape(h, stride, BPP[name])
            if name in ("UYVY422",):
                v[:, :w, 1] = g
            elif name == "RGB565LE":
                k16 = ((g.astype(np.uint16) >> 3) << 11) | ((g.astype(np.uint16) >> 2) << 5) | (g.astype(np.uint16) >> 3)
                v[:, :w, 0] = k16 & 0xff; v[:, :w, 1] = k16 >> 8
            else:
                for c in range(3):
                    v[:, :w, c] = g
    pitch = max(len(b) for b in bufs)
    frames = np.zeros((batch, pitch), np.uint8)
    for k, b in enumerate(bufs):
        frames[k, :len(b)] = b
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30)
    got = cvb.canny_kht_process_batch_fmt(canny, kht, FMT[name], frames, w, h, stride, pitch)
    kern = oracle.gauss_kernel("orc", 5, 1.0)
    total = 0
    for k in range(batch):
        gray = np.ascontiguousarray(gray_restated(name, frames[k], w, h, stride))
        edges = oracle.edge_dete("orc", oracle.convlt1("orc", "8u32f8u", gray, kern, kern), "canny", 59.0, 119.0, 3)
        want, _ = oracle.hough_kht("orc", edges, 1.0, 1.0, 30)
        assert len(got[k]) == len(want)
        for key in ("rho", "theta", "strength"):
            np.testing.assert_array_equal(got[k][key], want[key])
        total += len(want)
    assert total > 0
