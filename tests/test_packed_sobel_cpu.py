"""The packed 16-bit Sobel of the TMA fast paths (compv_b200/csrc/canny_fast.cuh, stage S3 / sobel_fast_kernel) restated with numpy uint32 arithmetic and
checked against the plain definition, on EVERY extreme neighbourhood (all 2^18 patterns of 0 / 255 in the 3 x 6 bytes a lane's four pixels read) and on random bytes:
the bias scheme must never let a borrow cross the 16-bit halves.  CPU only; the GPU parity tests check the kernels themselves."""
import numpy as np

U = np.uint32


def byte_perm(x, y, sel):
    """__byte_perm for selector nibbles 0..7 (no sign replication)."""
    b = [(x >> U(8 * i)) & U(0xff) for i in range(4)] + [(y >> U(8 * i)) & U(0xff) for i in range(4)]
    out = np.zeros_like(x)
    for i in range(4):
        out |= b[(sel >> (4 * i)) & 7] << U(8 * i)
    return out


def vmaxu2(a, b):
    lo = np.maximum(a & U(0xffff), b & U(0xffff))
    hi = np.maximum(a >> U(16), b >> U(16))
    return lo | (hi << U(16))


def load_row(wl, wc, wr):
    zero = np.zeros_like(wc)
    B = byte_perm(wc, zero, 0x4240)   # (p0, p2)
    Cc = byte_perm(wc, zero, 0x4341)  # (p1, p3)
    A = byte_perm(wl, Cc, 0x5453)     # (p-1, p1)
    D = byte_perm(B, wr, 0x1412)      # (p2, p4)
    hs = (A + U(2) * B + Cc, B + U(2) * Cc + D)
    hd = (Cc + U(0x01000100) - A, D + U(0x01000100) - B)
    return hs, hd


def packed_g(rows):
    """rows: three (wl, wc, wr) word triples (image rows y-1, y, y+1) -> g of the word's four pixels, as in the kernels."""
    (hsa, hda), (hsb, hdb), (hsc, hdc) = [load_row(*r) for r in rows]
    g = []
    for h in range(2):
        gx = hda[h] + U(2) * hdb[h] + hdc[h]
        gy = hsc[h] + U(0x04000400) - hsa[h]
        ax = vmaxu2(gx, U(0x08000800) - gx)
        ay = vmaxu2(gy, U(0x08000800) - gy)
        g.append(ax + ay - U(0x08000800))
    # g[0] = (g0 | g2 << 16), g[1] = (g1 | g3 << 16)
    return np.stack([g[0] & U(0xffff), g[1] & U(0xffff), g[0] >> U(16), g[1] >> U(16)], -1).astype(np.int64)


def plain_g(px):
    """px: (..., 3, 6) bytes = columns x-1 .. x+4 of three rows -> |gx| + |gy| of the four middle pixels."""
    p = px.astype(np.int64)
    out = []
    for i in range(1, 5):
        gx = (p[..., 0, i + 1] - p[..., 0, i - 1]) + 2 * (p[..., 1, i + 1] - p[..., 1, i - 1]) + (p[..., 2, i + 1] - p[..., 2, i - 1])
        gy = (p[..., 2, i - 1] + 2 * p[..., 2, i] + p[..., 2, i + 1]) - (p[..., 0, i - 1] + 2 * p[..., 0, i] + p[..., 0, i + 1])
        out.append(np.abs(gx) + np.abs(gy))
    return np.stack(out, -1)


def words(px):
    """(..., 3, 6) bytes -> per row (wl, wc, wr): wl's top byte is column x-1, wc the four pixels, wr's low byte column x+4 (other bytes arbitrary)."""
    rows = []
    rng = np.random.default_rng(7)
    for r in range(3):
        junk_l = rng.integers(0, 1 << 24, px.shape[:-2], dtype=np.uint32)
        junk_r = rng.integers(0, 1 << 24, px.shape[:-2], dtype=np.uint32) << U(8)
        p = px[..., r, :].astype(np.uint32)
        wl = junk_l | (p[..., 0] << U(24))
        wc = p[..., 1] | (p[..., 2] << U(8)) | (p[..., 3] << U(16)) | (p[..., 4] << U(24))
        wr = junk_r | p[..., 5]
        rows.append((wl, wc, wr))
    return rows


def test_packed_sobel_all_extreme_neighbourhoods():
    n = 1 << 18
    bits = (np.arange(n, dtype=np.uint32)[:, None] >> np.arange(18, dtype=np.uint32)[None, :]) & U(1)
    px = (bits * U(255)).astype(np.uint8).reshape(n, 3, 6)
    np.testing.assert_array_equal(packed_g(words(px)), plain_g(px))


def test_packed_sobel_random_bytes():
    px = np.random.default_rng(11).integers(0, 256, (400000, 3, 6), dtype=np.uint8)
    np.testing.assert_array_equal(packed_g(words(px)), plain_g(px))


def test_candidate_flag_by_biased_add():
    """canny_fast.cuh S3: adding 0x8000 - (tLow + 1) to both 16-bit halves sets a half's top bit exactly when its g > tLow (g <= 2040: nothing carries into the other half)."""
    g = np.arange(0, 2041, dtype=np.uint32)
    g0, g2 = np.meshgrid(g, g[::13], indexing="ij")
    pair = g0 | (g2 << U(16))
    for tlow in [1, 2, 59, 119, 2039, 2040, 2041, 32766, 32767, 65535]:
        add = U((0x8000 - (min(tlow, 0x7ffe) + 1)) * 0x10001)
        f = pair + add
        np.testing.assert_array_equal((f & U(0x8000)) != 0, g0 > tlow)
        np.testing.assert_array_equal((f & U(0x80000000)) != 0, g2 > tlow)
