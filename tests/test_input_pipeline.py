"""The on-device input pipeline (dopt_b200/csrc/input.cu) against the reference's host loops (oracle/input_ref.py:
cifar.d:50-55, imagetransformer.d:45-138).  Byte / index work: bit-exact.

CPU part: the literal restatement of ImageTransformer.getBatch against an independent formulation (np.pad symmetric + crop +
flips) and the closed-form gather the kernel implements.  GPU part: the kernel through the C ABI against the oracle on
identical draws, CIFAR- and SINS-shaped batches, odd widths, maximum jitter, and the device sampler's distributions."""
import numpy as np
import pytest

from oracle import input_ref as IR

F = np.float32


def draws_for(rng, n, jx, jy, flip_x=True, flip_y=True):
    d = np.zeros((n, 4), np.int32)
    d[:, 0] = rng.randint(0, 2 * jx, n) if jx else 0       # uniform(0, 2*jitter): upper bound exclusive
    d[:, 1] = rng.randint(0, 2 * jy, n) if jy else 0
    d[:, 2] = rng.randint(0, 2, n) if flip_x else 0
    d[:, 3] = rng.randint(0, 2, n) if flip_y else 0
    return d


def independent(batch, jx, jy, draws):
    out = np.empty_like(batch)
    N, C, H, W = batch.shape
    for n in range(N):
        xo, yo, fx, fy = draws[n]
        img = batch[n]
        if jx or jy:
            p = np.pad(img, ((0, 0), (jy, jy), (jx, jx)), mode="symmetric")
            img = p[:, yo:yo + H, xo:xo + W]
        if fx:
            img = img[:, :, ::-1]
        if fy:
            img = img[:, ::-1, :]
        out[n] = img
    return out


def closed_form(batch, jx, jy, draws):
    """out[c, y, x] = src[c, R_H(fy(y) + yOff - jy), R_W(fx(x) + xOff - jx)] -- what input.cu computes."""
    N, C, H, W = batch.shape

    def refl(i, n):
        return np.where(i < 0, -1 - i, np.where(i >= n, 2 * n - 1 - i, i))
    out = np.empty_like(batch)
    ys, xs = np.arange(H), np.arange(W)
    for n in range(N):
        xo, yo, fx, fy = (int(v) for v in draws[n])
        sy = refl((H - 1 - ys if fy else ys) + yo - jy, H)
        sx = refl((W - 1 - xs if fx else xs) + xo - jx, W)
        out[n] = batch[n][:, sy][:, :, sx]
    return out


@pytest.mark.parametrize("shape,jx,jy", [((5, 3, 32, 32), 4, 4), ((3, 1, 7, 9), 2, 3), ((2, 3, 6, 5), 5, 6),
                                         ((4, 2, 8, 8), 0, 0), ((2, 3, 96, 96), 12, 12), ((3, 2, 5, 4), 1, 1)])
def test_reference_loops_equal_symmetric_pad_crop_flip(shape, jx, jy):
    rng = np.random.RandomState(1)
    batch = rng.randn(*shape).astype(F)
    d = draws_for(rng, shape[0], jx, jy)
    want = IR.image_transform(batch, jx, jy, d)
    np.testing.assert_array_equal(want, independent(batch, jx, jy, d))
    np.testing.assert_array_equal(want, closed_form(batch, jx, jy, d))


def test_every_offset_and_flip_combination():
    rng = np.random.RandomState(2)
    batch = rng.randn(1, 2, 6, 5).astype(F)
    jx, jy = 2, 3
    for xo in range(2 * jx):
        for yo in range(2 * jy):
            for fx in (0, 1):
                for fy in (0, 1):
                    d = np.array([[xo, yo, fx, fy]], np.int32)
                    np.testing.assert_array_equal(IR.image_transform(batch, jx, jy, d), closed_form(batch, jx, jy, d))


def test_centre_crop_without_flip_is_the_identity_and_input_is_not_modified():
    rng = np.random.RandomState(3)
    batch = rng.randn(2, 3, 8, 8).astype(F)
    keep = batch.copy()
    d = np.array([[4, 4, 0, 0]] * 2, np.int32)
    np.testing.assert_array_equal(IR.image_transform(batch, 4, 4, d), keep)
    np.testing.assert_array_equal(batch, keep)


def test_normalisation_and_one_hot():
    raw = np.arange(256, dtype=np.uint8)
    v = IR.normalise_u8(raw)
    assert v[0] == -1.0 and v[128] == 0.0 and v[255] == F(255) / F(128) - F(1)
    assert np.array_equal(v, (raw.astype(np.float64) / 128.0 - 1.0).astype(F))     # exact in fp32
    oh = IR.one_hot(np.array([3, 0, 9], np.uint8), 10)
    assert oh.shape == (3, 10) and oh.sum() == 3 and oh[0, 3] == 1 and oh[1, 0] == 1 and oh[2, 9] == 1


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the kernel through the C ABI
# ---------------------------------------------------------------------------------------------------------------------
GPU_CASES = [((128, 3, 32, 32), 4, 4), ((50, 3, 96, 96), 12, 12), ((3, 1, 7, 9), 2, 3), ((2, 3, 6, 5), 5, 6),
             ((7, 2, 8, 8), 0, 0), ((1, 1, 1, 1), 1, 1), ((4, 3, 28, 28), 0, 3)]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,jx,jy", GPU_CASES)
def test_image_transform_kernel_is_bit_exact(shape, jx, jy):
    import torch
    import dopt_b200 as db
    rng = np.random.RandomState(4)
    raw = rng.randint(0, 256, shape).astype(np.uint8)
    d = draws_for(rng, shape[0], jx, jy)
    want = IR.image_transform(IR.normalise_u8(raw), jx, jy, d)
    got = db.image_transform(torch.from_numpy(raw).cuda(), jx, jy, torch.from_numpy(d).cuda())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # float source (what ImageTransformer itself is handed)
    f = rng.randn(*shape).astype(F)
    got = db.image_transform(torch.from_numpy(f).cuda(), jx, jy, torch.from_numpy(d).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), IR.image_transform(f, jx, jy, d))
    # no table: plain normalisation
    got = db.image_transform(torch.from_numpy(raw).cuda(), jx, jy, None)
    np.testing.assert_array_equal(got.cpu().numpy(), IR.normalise_u8(raw))


@pytest.mark.gpu
def test_one_hot_and_errors():
    import torch
    import dopt_b200 as db
    labels = np.random.RandomState(5).randint(0, 100, 128).astype(np.uint8)
    got = db.one_hot(torch.from_numpy(labels).cuda(), 100)
    np.testing.assert_array_equal(got.cpu().numpy(), IR.one_hot(labels, 100))
    x = torch.zeros((1, 1, 4, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(db.DoptError):
        db.image_transform(x, 5, 0, None)            # jitter larger than the image


@pytest.mark.gpu
def test_device_sampler_follows_the_reference_distributions():
    import torch
    import dopt_b200 as db
    n, jx, jy = 1 << 16, 4, 6
    a = db.jitter_sample(n, jx, jy, True, True, seed=7, call=0).cpu().numpy()
    b = db.jitter_sample(n, jx, jy, True, True, seed=7, call=0).cpu().numpy()
    c = db.jitter_sample(n, jx, jy, True, True, seed=7, call=1).cpu().numpy()
    np.testing.assert_array_equal(a, b)                                   # counter-based: reproducible
    assert not np.array_equal(a, c)                                       # a new call draws new numbers
    assert a[:, 0].min() == 0 and a[:, 0].max() == 2 * jx - 1             # uniform(0, 2*jitter), upper bound exclusive
    assert a[:, 1].min() == 0 and a[:, 1].max() == 2 * jy - 1
    for col, k in ((0, 2 * jx), (1, 2 * jy), (2, 2), (3, 2)):
        freq = np.bincount(a[:, col], minlength=k) / float(n)
        assert np.abs(freq - 1.0 / k).max() < 0.01, (col, freq)
    off = db.jitter_sample(16, 0, 0, False, False, seed=7).cpu().numpy()
    assert not off.any()
    # end to end: sampled draws through the kernel equal the oracle on the same draws
    raw = np.random.RandomState(6).randint(0, 256, (64, 3, 32, 32)).astype(np.uint8)
    d = db.jitter_sample(64, 4, 4, True, False, seed=11)
    got = db.image_transform(torch.from_numpy(raw).cuda(), 4, 4, d)
    np.testing.assert_array_equal(got.cpu().numpy(), IR.image_transform(IR.normalise_u8(raw), 4, 4, d.cpu().numpy()))
