"""CPU oracle for the training-time augmentation of the face / parsing-map pairs.  TEST INFRASTRUCTURE ONLY.

The reference draws one angle and three factors per sample and applies, to the LR and the HR image alike
(/root/reference/helen_loader.py:75-101; the parsing map is only rotated, :103-104):

    img = img.rotate(angle)                                  # PIL: NEAREST, no expand, zero fill
    img = ImageEnhance.Contrast(img).enhance(contrast)       # all three enhancers are Contrast in the reference
    img = ImageEnhance.Contrast(img).enhance(brightness)     # (:87-91, :97-101 construct Contrast for "brightness"
    img = ImageEnhance.Contrast(img).enhance(sharpness)      #  and "sharpness" as well)

and rotates the landmark coordinates by the opposite angle about the image centre (:110-113, rotate_matrix :272-275).
The arithmetic lives in Pillow (third party, unpinned in the reference; 12.2.0 in this image).  This file restates the
published algorithms in numpy / Python:

  * Image.rotate -> Image.transform(AFFINE, NEAREST): matrix from cos / sin rounded to 15 decimals, centre (w/2, h/2);
    ImagingTransformAffine's 16.16 fixed-point path: FIX(v) = floor(v * 65536 + 0.5), start values include the half
    pixel offset, source index = accumulator >> 16, pixels that fall outside stay 0;
  * ImageEnhance.Contrast: degenerate image = int(mean(L) + 0.5) with L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16;
    Image.blend(degenerate, image, f) in float32: t = in1 + f * (in2 - in1); 0 <= f <= 1: (uint8) t (truncation),
    otherwise clipped to [0, 255] first; f == 0 / f == 1 return copies.

Pin: ``oracle/make_golden.py augment`` compares these functions with Pillow on random images (bit-exact) and stores the
vectors under ``tests/golden/augment.npz``.
"""
import math

import numpy as np


def _fix(v):
    t = v * 65536.0 + 0.5
    return int(t) if t >= 0.0 else int(math.floor(t))


def rotate_coeffs(h, w, angle_deg):
    """The six 16.16 fixed-point coefficients (a0..a5) of Pillow's affine_fixed for Image.rotate(angle_deg)."""
    angle = angle_deg % 360.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2, h / 2
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return np.array([_fix(m[0]), _fix(m[1]), _fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
                     _fix(m[3]), _fix(m[4]), _fix(m[5] + m[3] * 0.5 + m[4] * 0.5)], np.int32)


def rotate_u8(img, angle_deg):
    """img uint8 [h][w][c] -> Image.rotate(angle_deg) (general path; the reference draws angles in (-10, 10))."""
    h, w = img.shape[:2]
    if angle_deg % 360.0 == 0:
        return img.copy()
    a0, a1, a2, a3, a4, a5 = (int(v) for v in rotate_coeffs(h, w, angle_deg))
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def luma_mean(img):
    if img.ndim == 3 and img.shape[2] == 3:
        r, g, b = (img[..., i].astype(np.int64) for i in range(3))
        lum = (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16
    else:
        lum = img.reshape(img.shape[0], img.shape[1]).astype(np.int64)
    return int(float(lum.sum()) / lum.size + 0.5)


def contrast_u8(img, factor):
    """ImageEnhance.Contrast(img).enhance(factor) for an RGB or single-channel uint8 image."""
    f = np.float32(factor)
    mean = luma_mean(img)
    if f == np.float32(0.0):
        return np.full_like(img, mean)
    if f == np.float32(1.0):
        return img.copy()
    in1 = np.float32(mean)
    t = (f * (img.astype(np.float32) - in1)).astype(np.float32)
    t = (in1 + t).astype(np.float32)
    if np.float32(0.0) <= f <= np.float32(1.0):
        return t.astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255.0, 255, t.astype(np.int64))).astype(np.uint8)


def augment_u8(img, angle_deg, factors=()):
    """helen_loader.py:75-101 for one image: rotate, then one Contrast enhancement per factor, in order."""
    out = rotate_u8(img, angle_deg)
    for f in factors:
        out = contrast_u8(out, f)
    return out


def rotate_landmarks(landmarks, angle_deg, centre):
    """helen_loader.py:110-113: (x, y) rows rotated by -angle about `centre` with rotate_matrix (:272-275)."""
    a = -angle_deg / 180.0 * math.pi
    m = np.array([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
    return (np.asarray(landmarks, np.float64) - centre) @ m.T + centre
