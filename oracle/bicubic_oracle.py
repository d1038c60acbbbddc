"""CPU oracle for the offline LR->input bicubic upsample.  TEST INFRASTRUCTURE ONLY.

The reference resizes every low-resolution face with ``img.resize((W, H), Image.BICUBIC)``
(/root/reference/bicubic_interpolation.py:188, /root/reference/SUPER_RESOLUTION/FHN_loader.py:66).  The arithmetic
lives in Pillow (third-party, unpinned in the reference; 12.2.0 in this image), function ``ImagingResample`` with
the 8-bit fixed-point path.  This file restates that published algorithm in numpy:

  * separable: horizontal pass first, then vertical; the intermediate image is uint8;
  * Keys cubic, a = -0.5, support 2 (x filterscale = max(in/out, 1));
  * per output index: window [xmin, xmax) clipped to the image, weights renormalised to sum 1;
  * weights -> int(w * 2**22 +- 0.5); accumulator starts at 2**21; result = clip(acc >> 22, 0, 255).

Pin: ``oracle/make_golden.py`` compares this function with Pillow on random images (bit-exact) and stores the
vectors under ``tests/golden/bicubic_*.npz``.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _keys(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeff_table(in_size, out_size):
    """Returns (xmin[out], count[out], kk[out, ksize] int32) exactly as Pillow's precompute_coeffs + normalize_coeffs_8bpc."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.array([_keys((x + lo - center + 0.5) * ss) for x in range(n)], np.float64)
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if w[x] < 0 else int(v + 0.5)
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass_axis(img, out_size, axis):
    """img uint8 [..., H, W, C]; resample ``axis`` (-3 = rows, -2 = columns)."""
    in_size = img.shape[axis]
    xmin, cnt, kk = coeff_table(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(cnt[xx]):
            acc += src[xmin[xx] + t] * int(kk[xx, t])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def bicubic_u8(img_hwc, out_h, out_w):
    """uint8 [H, W, C] (or [N, H, W, C]) -> uint8 [out_h, out_w, C]; bit-exact with PIL Image.BICUBIC."""
    img = np.asarray(img_hwc, np.uint8)
    tmp = _pass_axis(img, out_w, -2) if img.shape[-2] != out_w else img
    return _pass_axis(tmp, out_h, -3) if img.shape[-3] != out_h else tmp


def normalise_to_input(u8_nhwc):
    """helen_loader.py:53-58 style tensor conversion: /255 then (x - 0.5) / 0.5, NCHW float32."""
    x = u8_nhwc.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(np.moveaxis(x, -1, -3))


def landmark_heatmap(landmarks, height, width, sigma=1.3):
    """Restatement of HelenLoader.generate_hm / gaussian_k (/root/reference/helen_loader.py:118-143): the sum over the
    landmarks of exp(-((x - x0)^2 + (y - y0)^2) / (2 sigma^2)), Gaussians in float64, accumulated in place into a
    float32 map (i.e. rounded to float32 after every landmark).  landmarks: [K, 2] (x, y).  TEST INFRASTRUCTURE ONLY."""
    x = np.arange(0, width, 1, float)
    y = np.arange(0, height, 1, float)[:, np.newaxis]
    hm = np.zeros((height, width), dtype=np.float32)
    for x0, y0 in np.asarray(landmarks, dtype=np.float64):
        hm += np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
    return hm
