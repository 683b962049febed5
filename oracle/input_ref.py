"""CPU oracle for the input pipeline -- TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).

Literal restatements of the reference's host loops, kept in the reference's own shape (a reflect-padded scratch image,
then a crop, then in-place flips) so that they check the closed-form gather the CUDA kernel uses rather than repeat it:

  normalise_u8      nnet/source/dopt/nnet/data/cifar.d:50      `cast(T)x / 128.0f - 1.0f`
  one_hot           nnet/source/dopt/nnet/data/cifar.d:52-55
  image_transform   nnet/source/dopt/nnet/data/imagetransformer.d:45-138 (ImageTransformer.getBatch), with the random draws
                    (`uniform(0, 2*jitter)` twice, `uniform(0.0f, 1.0f) < 0.5f` per enabled flip, :101-102,118,126) passed
                    in per image as (x_off, y_off, flip_x, flip_y) instead of drawn from std.random.

Parity is unpinned by the reference (it has no test for ImageTransformer); the restatement is cross-checked against
numpy's `np.pad(mode="symmetric")` + slicing in tests/test_input_pipeline.py.
"""
import numpy as np


def normalise_u8(raw):
    return (raw.astype(np.float32) / np.float32(128.0) - np.float32(1.0)).astype(np.float32)


def one_hot(labels, classes):
    out = np.zeros((len(labels), classes), np.float32)
    for i, l in enumerate(labels):
        out[i, int(l)] = 1.0
    return out


def image_transform(batch, jitter_x, jitter_y, draws):
    """batch: float32 [N, C, H, W]; draws: int [N, 4] = (x_off, y_off, flip_x, flip_y) per image.  Returns a new array."""
    batch = np.array(batch, dtype=np.float32, copy=True)
    N, C, H, W = batch.shape
    pw, ph = W + 2 * jitter_x, H + 2 * jitter_y
    padded = np.zeros(C * ph * pw, np.float32)                                   # mPadded, :17
    for n in range(N):
        img = batch[n].reshape(-1)                                               # one chunk of volume[0] floats, :57
        x_off, y_off, flip_x, flip_y = (int(v) for v in draws[n])
        if jitter_x != 0 or jitter_y != 0:
            for c in range(C):
                for y in range(H):
                    o = c * ph * pw + (y + jitter_y) * pw
                    padded[o + jitter_x:o + jitter_x + W] = img[c * H * W + y * W:c * H * W + (y + 1) * W]   # :63-71
                    if jitter_x != 0:
                        padded[o:o + jitter_x] = padded[o + jitter_x:o + 2 * jitter_x][::-1].copy()         # :75-77
                        o2 = o + W
                        padded[o2 + jitter_x:o2 + 2 * jitter_x] = padded[o2:o2 + jitter_x][::-1].copy()      # :79-81
                for y in range(jitter_y):
                    o = c * pw * ph
                    padded[o + y * pw:o + (y + 1) * pw] = \
                        padded[o + (2 * jitter_y - y - 1) * pw:o + (2 * jitter_y - y) * pw].copy()           # :89-91
                    padded[o + (ph - y - 1) * pw:o + (ph - y) * pw] = \
                        padded[o + (ph - 2 * jitter_y + y) * pw:o + (ph - 2 * jitter_y + y + 1) * pw].copy()  # :93-95
            p3 = padded.reshape(C, ph, pw)
            img[:] = p3[:, y_off:y_off + H, x_off:x_off + W].reshape(-1)          # crop, :104-115
        if flip_x:
            rows = img.reshape(-1, W)
            rows[:] = rows[:, ::-1].copy()                                        # every row reversed, :118-124
        if flip_y:
            maps = img.reshape(C, H, W)
            maps[:] = maps[:, ::-1, :].copy()                                     # every column reversed, :126-137
    return batch
