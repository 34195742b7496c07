"""TEST INFRASTRUCTURE (like everything under oracle/): a numpy restatement of OpenCV's 8-bit fixed-point cubic
`cv2.warpAffine` (imgproc/src/imgwarp.cpp: WarpAffineInvoker, remapBicubic, initInterTab2D, interpolateCubic) — the
algorithm csrc/patches.cu implements on the GPU for the patch extraction of utils/library.py:84-110.  OpenCV is a
third-party dependency of the reference (opencv-python, any 4.x: the 8-bit path is bit-exact across versions by design);
tests/test_patches_cpu.py pins this restatement to the installed cv2 on random warps, and the C library's weight table and
the GPU kernel to this restatement.  Slow (pure Python per pixel): small cases only."""
import numpy as np

INTER_BITS, TAB, COEF_BITS = 5, 32, 15
SCALE = 1 << COEF_BITS


def cubic_coeffs(x):
    """interpolateCubic: float arithmetic, A = -0.75."""
    x = np.float32(x)
    a, one = np.float32(-0.75), np.float32(1)
    c0 = ((a * (x + one) - np.float32(5) * a) * (x + one) + np.float32(8) * a) * (x + one) - np.float32(4) * a
    c1 = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + one
    c2 = ((a + np.float32(2)) * (one - x) - (a + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def bicubic_table():
    """[32 fy][32 fx][4][4] integer weights at scale 2^15, every set summing to exactly 2^15 (initInterTab2D)."""
    t1 = np.stack([cubic_coeffs(np.float32(i) * np.float32(1.0 / TAB)) for i in range(TAB)])
    itab = np.zeros((TAB, TAB, 4, 4), dtype=np.int32)
    for i in range(TAB):
        for j in range(TAB):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32)
            it = np.clip(np.rint((v * np.float32(SCALE)).astype(np.float32)).astype(np.int64), -32768, 32767)
            isum = int(it.sum())
            if isum != SCALE:
                diff = isum - SCALE
                big = small = (2, 2)
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        if it[k1, k2] < it[small]:
                            small = (k1, k2)
                        elif it[k1, k2] > it[big]:
                            big = (k1, k2)
                if diff < 0:
                    it[big] -= diff
                else:
                    it[small] -= diff
            itab[i, j] = it
    return itab


def invert_affine(m):
    """cv::warpAffine without WARP_INVERSE_MAP: the 2 x 3 matrix in double, inverted in this order of operations."""
    m = np.asarray(m, dtype=np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_affine_cubic_u8(src, m, dsize, itab=None):
    """cv2.warpAffine(src uint8 (H, W[, C]), m, dsize, flags=INTER_CUBIC, borderMode=BORDER_CONSTANT, borderValue=0)."""
    itab = bicubic_table() if itab is None else itab
    m = invert_affine(m)
    wd, hd = dsize
    h, w = src.shape[:2]
    cn = src.shape[2] if src.ndim == 3 else 1
    s = src.reshape(h, w, cn).astype(np.int64)
    ab, abs_ = 10, 1 << 10
    rd = abs_ // TAB // 2
    xs = np.arange(wd)
    adelta = np.rint(m[0] * xs * abs_).astype(np.int64)
    bdelta = np.rint(m[3] * xs * abs_).astype(np.int64)
    out = np.zeros((hd, wd, cn), dtype=np.uint8)
    for y in range(hd):
        x0 = int(np.rint((m[1] * y + m[2]) * abs_)) + rd
        y0 = int(np.rint((m[4] * y + m[5]) * abs_)) + rd
        xx = (x0 + adelta) >> (ab - INTER_BITS)
        yy = (y0 + bdelta) >> (ab - INTER_BITS)
        sx, sy = (xx >> INTER_BITS) - 1, (yy >> INTER_BITS) - 1
        fx, fy = xx & (TAB - 1), yy & (TAB - 1)
        for x in range(wd):
            wt = itab[fy[x], fx[x]]
            acc = np.zeros(cn, dtype=np.int64)
            for k1 in range(4):
                r = sy[x] + k1
                if r < 0 or r >= h:
                    continue
                for k2 in range(4):
                    c = sx[x] + k2
                    if c < 0 or c >= w:
                        continue
                    acc += s[r, c] * int(wt[k1, k2])
            out[y, x] = np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255)
    return out if src.ndim == 3 else out[:, :, 0]
