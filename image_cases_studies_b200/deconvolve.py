"""Pyramid driver: drop-in for the reference's ``deconvolve.py`` (``deblur_module``, ``pad_image``, ``build_pyramid``).

Same arguments and flow as ``deconvolve.py:65-368``: gamma removal, odd padding, mask window, coarse-to-fine pyramid,
a blind pass on the mask crop followed by a non-blind pass on the full frame at every level, clipping, gamma,
16-bit TIFF.  Every solver call goes to the B200 path (``lib.deconvolution``); the glue around it is host-side
numpy, as in the reference.

Parity note: the reference's driver cannot be executed in the build container (scikit-image and
matplotlib are absent), so this file is **unpinned by reference execution**.  Its one numerical dependency outside
the solver, ``skimage.transform.resize(order=3, mode="edge", preserve_range=True)``, is restated with
``scipy.ndimage.zoom(order=3, mode="nearest", grid_mode=True)`` (what scikit-image calls underneath), without
anti-aliasing (the default of the scikit-image releases contemporary with the reference).
"""
from __future__ import annotations

import numpy as np

from .lib import utils


def pad_image(image, pad, mode="edge"):
    """deconvolve.py:24-37: pad the two spatial axes of an RGB image, return contiguous float32.  ``pad`` is handed to
    ``np.pad`` per channel exactly as the reference does, so an int, a (before, after) pair and a per-axis pair of pairs
    all mean what they mean there."""
    planes = [np.pad(image[..., c], pad, mode=mode) for c in range(image.shape[2])]
    return np.ascontiguousarray(np.stack(planes, axis=2), np.float32)


def build_pyramid(psf_size, lambd):
    """deconvolve.py:40-60: image scales 1, 1/sqrt2, ... and odd kernel sizes down to 3."""
    images = [1.]
    kernels = [psf_size]
    while kernels[-1] > 3:
        kernels.append(int(np.ceil(kernels[-1] / np.sqrt(2))))
        images.append(images[-1] / np.sqrt(2))
        if kernels[-1] % 2 == 0:
            kernels[-1] -= 1
        if kernels[-1] < 3:
            kernels[-1] = 3
    return images, kernels


def resize(image, shape):
    """Bicubic-spline resize of an (H, W, C) array to ``shape`` (restates skimage.transform.resize, see module doc)."""
    from scipy import ndimage
    image = np.asarray(image)
    zoom = (shape[0] / image.shape[0], shape[1] / image.shape[1])
    out = np.empty(tuple(shape), dtype=np.float64)
    for c in range(shape[2]):
        out[..., c] = ndimage.zoom(image[..., c].astype(np.float64), zoom, order=3, mode="nearest", grid_mode=True)
    return out


_QUALITY_STEP = {"normal": 1e-3, "high": 5e-4, "veryhigh": 1e-4, "low": 5e-3}     # deconvolve.py:106-113


@utils.timeit
def deblur_module(pic, filename, dest_path, blur_width, confidence=10, tolerance=1, quality="normal", bits=8,
                  mask=None, display=True, blur="static", preview=False, p=1, order=2, norm=1, priority=0, mask_size=255,
                  iterations=200, refocus=False, solver=None, save=True):
    """API to call the deblurring process (deconvolve.py:65-368).  Returns the 16-bit-scaled result (the reference
    only writes it to ``<dest_path>/<filename>.tif``; pass ``save=False`` to skip the file)."""
    dc = solver
    if dc is None:
        from .lib import deconvolution as dc
    pic = np.ascontiguousarray(pic, dtype=np.float32)
    pic = pad_image(pic, (1, 1)).astype(np.float32)                               # :94
    samples = 2 ** bits - 1
    pic = pic / samples                                                           # :100
    pic = pic ** (1 / 2.2)                                                        # :103
    step = _QUALITY_STEP[quality]
    if blur_width < 3:
        raise ValueError("The blur width should be at least 3 pixels.")
    elif blur_width % 2 == 0:
        raise ValueError("The blur width should be odd. You can use %i." % (blur_width + 1))
    MK = blur_width
    M, N = pic.shape[0], pic.shape[1]
    if mask is None:
        mask = [M // 2, N // 2]
    top, bottom = mask[0] - mask_size // 2, mask[0] + mask_size // 2               # :138-141
    left, right = mask[1] - mask_size // 2, mask[1] + mask_size // 2
    if not (top > 0 and bottom < M and left > 0 and right < N):
        raise ValueError("The mask is outside the picture boundaries. Move its center inside or reduce the blur size.")
    correlation = {"static": False, "motion": True}[blur]                         # :154-157
    tolerance /= 100.
    odd_vert = odd_hor = False
    if pic.shape[0] % 2 == 0:                                                      # :164-175
        pic = np.ascontiguousarray(np.pad(pic, ((1, 0), (0, 0), (0, 0)), mode="edge"), np.float32)
        odd_vert = True
    if pic.shape[1] % 2 == 0:
        pic = np.ascontiguousarray(np.pad(pic, ((0, 0), (1, 0), (0, 0)), mode="edge"), np.float32)
        odd_hor = True
    psf = utils.stack3(utils.uniform_kernel(blur_width))                          # :178-179
    images, kernels = build_pyramid(blur_width, confidence)

    deblured_image = pic.copy()
    for case in ("blind", "non-blind"):                                            # :193
        deblured_image = pic.copy()
        lambd = confidence * 1000
        for i, k in zip(reversed(images), reversed(kernels)):                      # :204
            t_top, t_bottom = int(i * top), int(i * bottom)
            t_left, t_right = int(i * left), int(i * right)
            if (t_bottom - t_top) % 2 == 0:                                        # :215-221 odd, square-ish mask
                if (t_bottom - t_top) < (t_right - t_left):
                    t_bottom += 1
                elif (t_bottom - t_top) > (t_right - t_left):
                    t_top += 1
                else:
                    t_top -= 1
            if (t_right - t_left) % 2 == 0:                                        # :223-229 (the reference's second test
                if (t_bottom - t_top) < (t_right - t_left):                        #  compares a quantity with itself, so
                    t_left += 1                                                    #  only these two branches can run)
                else:
                    t_right += 1
            t_width, t_height = int(np.floor(i * N)), int(np.floor(i * M))
            t_width += (t_width % 2 == 0)
            t_height += (t_height % 2 == 0)
            shape = (t_height, t_width, 3)
            blurry = resize(pic, shape).astype(np.float32)                         # :245-246
            deblured_image = resize(deblured_image, shape).astype(np.float32)
            if case == "blind":
                psf_copy = np.ascontiguousarray(resize(psf, (k, k, 3)).astype(np.float32))
                dc.normalize_kernel(psf_copy, k)                                   # :249-250
            else:
                psf_copy = psf.copy()
                k = kernels[0]
            blurry = pad_image(blurry, (1, 1))                                     # :256-257
            deblured_image = pad_image(deblured_image, (1, 1))
            pad = int(np.floor(k / 2))
            tol = tolerance if i == 1. else 0                                      # :270-273
            win = (pad + 1, t_bottom - t_top - pad - 1, pad + 1, t_bottom - t_top - pad - 1)
            if case == "blind" or preview:
                img_v = blurry[t_top - 1:t_bottom + 1, t_left - 1:t_right + 1, ...]
                u_v = deblured_image[t_top - pad - 1:t_bottom + pad + 1, t_left - pad - 1:t_right + pad + 1, ...]
                out = dc.richardson_lucy_MM(img_v, u_v, psf_copy, *win, 0 if case == "blind" else tol,
                                            t_bottom - t_top + 2, t_right - t_left + 2, 3, k, iterations, step, lambd,
                                            blind=(case == "blind"), p=p, correlation=correlation if case == "blind" else False,
                                            order=order, norm=2, priority=0 if case == "blind" else priority, refocus=refocus)
                deblured_image[t_top - 1:t_bottom + 1, t_left - 1:t_right + 1, ...] = out    # :277-300
                if case == "blind":
                    psf = psf_copy.copy()                                          # :288
            else:
                deblured_image = pad_image(deblured_image, (pad, pad))             # :303
                out = dc.richardson_lucy_MM(blurry, deblured_image, psf_copy, *win, tol, t_height + 2, t_width + 2, 3, k,
                                            iterations, step, lambd, blind=False, p=p, order=order, norm=2,
                                            priority=priority, refocus=refocus)
                deblured_image[pad:-pad, pad:-pad, ...] = out
                deblured_image = deblured_image[pad:-pad, pad:-pad, ...]           # :316
            deblured_image = np.ascontiguousarray(deblured_image[1:-1, 1:-1, ...])  # :322-323

    deblured_image = np.clip(deblured_image, 0., 1.)                               # :346
    deblured_image = deblured_image ** 2.2                                         # :349
    deblured_image = deblured_image * (2 ** 16 - 1)                                # :352
    if preview:
        filename = filename + "-preview"
        deblured_image = deblured_image[top:bottom, left:right, ...]
    else:
        if odd_hor:
            deblured_image = deblured_image[:, 1:, ...]
        if odd_vert:
            deblured_image = deblured_image[1:, :, ...]
        deblured_image = deblured_image[1:-1, 1:-1, ...]                           # :366
    if save:
        utils.save(deblured_image, filename, dest_path)                            # :368
    deblur_module.last_psf = psf
    return deblured_image
