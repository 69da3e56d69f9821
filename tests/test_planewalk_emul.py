"""CPU model of the plane-walk convolution's GEMM decomposition (csrc/unet_tcz.cu), against a direct 3x3x3 convolution.

The CUDA kernel itself is tested on the GPU (tests/test_gpu_lcn_unet.py::test_conv_block_planewalk_kernel); this file pins
the ARITHMETIC PLAN it implements, in NumPy, so that the layout rules written in the kernel's header can be checked
without a GPU:
  * operands as fp16 hi / lo' images (x s = hi + lo' 2^-11), three product terms kept (hi.hi, hi.lo', lo'.hi);
  * B rows of one (K step, group): 48 blk + 24 term + 8 dz + co, blk 0 / 1 / 2 = x-tap dx 2 / 1 / 0; image 0 = (hi | lo')
    rows for A_hi, image 1 = (0 | hi) rows for A_lo';
  * K = (channel chunk, dy) taps in pairs, an odd tap count ends with (zero weights, last tap);
  * M tile = 8 y-rows x 16 z of one x-plane, rows 16 y + z; the MMA of input plane j accumulates onto the ring blocks of
    output planes j-1, j, j+1 (first contribution overwrites), a plane is read once after its third contribution;
  * the z shift-add: out[z] = c[dz=1][z] + c[dz=0][z-1] + c[dz=2][z+1] with zeros outside the tile.
Reference semantics: Conv3D(3, padding='same') of unet3d.py:117, per tile (zero padding at the tile border)."""
import numpy as np
import pytest

R = 5                       # ring slots per lane (TZ_R)
BLK = 48                    # accumulator columns of one output plane and group


def split_fp16(x, s):
    hi = (x * s).astype(np.float16)
    lo = ((x * s - hi.astype(np.float64)) * 2048.0).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def pack_images(w, g):
    """w: (3,3,3,cin,cout) keras kernel (kx,ky,kz,ci,co), group g of 8 output channels ->
    list over K steps of (image0, image1), each (2 K-halves, 144 rows, 8 channels) of fp16-valued float64, and 1/scale."""
    cin = w.shape[3]
    ntaps = 3 * (cin // 8)
    nsteps = (ntaps + 1) // 2
    scale = 2.0 ** (14 - np.frexp(np.abs(w).max())[1])
    steps = []
    for p in range(nsteps):
        img1, img2 = np.zeros((2, 144, 8)), np.zeros((2, 144, 8))
        for j in range(2):
            t = 2 * p + j
            if 2 * p + 1 >= ntaps:
                if j == 0:
                    continue                      # zero weights against the repeated tap
                t = ntaps - 1
            c, dy = divmod(t, 3)
            for blk in range(3):
                dx = 2 - blk
                for dz in range(3):
                    for col in range(8):
                        hi, lo = split_fp16(w[dx, dy, dz, c * 8:(c + 1) * 8, 8 * g + col], scale)
                        r = blk * BLK + dz * 8 + col
                        img1[j, r], img1[j, r + 24], img2[j, r + 24] = hi, lo, hi
        steps.append((img1, img2))
    return steps, 1.0 / scale


def a_views(plane_hi, plane_lo, p, ntaps, y0):
    """A operand of K step p for the M tile at rows y0..y0+7: (128 rows = 16 y + z, 2 K-halves, 8 channels).
    plane_*: (Y + 2 haloed rows, 16 z, cin) of one x-plane (row 0 = y -1)."""
    t1 = 2 * p + 1 if 2 * p + 1 < ntaps else ntaps - 1
    out = []
    for img in (plane_hi, plane_lo):
        halves = []
        for t in (t1 - 1, t1):
            c, dy = divmod(t, 3)
            halves.append(img[y0 + dy:y0 + dy + 8, :, c * 8:(c + 1) * 8].reshape(128, 8))
        out.append(np.stack(halves, 1))
    return out                                    # [A_hi, A_lo]


@pytest.mark.parametrize("cin,cout,X,Y", [(8, 8, 7, 16), (16, 16, 6, 8), (32, 8, 5, 16)])
def test_planewalk_plan_matches_direct_convolution(cin, cout, X, Y):
    rng = np.random.default_rng(cin * 100 + cout)
    Z = 16
    x = rng.normal(0, 1, (X, Y, Z, cin))
    w = rng.normal(0, 0.2, (3, 3, 3, cin, cout))
    # ---- direct 'same' convolution with zero padding
    xp = np.pad(x, ((1, 1), (1, 1), (1, 1), (0, 0)))
    want = np.zeros((X, Y, Z, cout))
    for dx in range(3):
        for dy in range(3):
            for dz in range(3):
                want += np.einsum("xyzc,co->xyzo", xp[dx:dx + X, dy:dy + Y, dz:dz + Z], w[dx, dy, dz])
    # ---- the kernel's plan
    s_in = 2.0 ** (13 - np.floor(np.log2(np.abs(x).max())))
    hi, lo = split_fp16(x, s_in)
    hi = np.pad(hi, ((1, 1), (1, 1), (0, 0), (0, 0)))           # haloed planes / rows (TMA zero fill); no z halo
    lo = np.pad(lo, ((1, 1), (1, 1), (0, 0), (0, 0)))
    ntaps = 3 * (cin // 8)
    got = np.zeros_like(want)
    for g in range(cout // 8):
        steps, inv_w = pack_images(w, g)
        for y0 in range(0, Y, 8):
            ring = np.full((R, 128, BLK), np.nan)               # NaN: a block must be overwritten before it is used
            for h in range(X + 2):                              # haloed input planes; newest output id = h
                d = np.zeros((128, 144))
                for p, (img1, img2) in enumerate(steps):
                    a_hi, a_lo = a_views(hi[h], lo[h], p, ntaps, y0)
                    d += np.einsum("rjk,jnk->rn", a_hi, img1) + np.einsum("rjk,jnk->rn", a_lo, img2)
                for blk in range(3):                            # output planes h-2 (accumulate), h-1, h (overwrite)
                    i = h - 2 + blk
                    if i < 0:
                        continue                                # pseudo ids in front of the segment
                    slot = i % R
                    part = d[:, blk * BLK:(blk + 1) * BLK]
                    ring[slot] = part if blk == 2 else ring[slot] + part
                i = h - 2                                       # complete after its third contribution: drained once
                if 0 <= i < X:
                    blkv = ring[i % R].reshape(8, 16, 2, 3, 8)  # (y, z, term, dz, co)
                    c = blkv[:, :, 0] + blkv[:, :, 1] / 2048.0  # hi.hi + 2^-11 (hi.lo' + lo'.hi)
                    out = c[:, :, 1].copy()
                    out[:, 1:] += c[:, :-1, 0]                  # tap dz = 0 comes from input row z - 1
                    out[:, :-1] += c[:, 1:, 2]                  # tap dz = 2 from input row z + 1
                    got[i, y0:y0 + 8, :, 8 * g:8 * g + 8] = out * inv_w / s_in
    assert np.isfinite(got).all()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 2e-6, err                                      # the dropped lo'.lo' term is 2^-22 relative
