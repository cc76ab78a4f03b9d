"""GPU triplet batcher against a restatement of MonoDataset.__getitem__ + default collation
(vo/dataset/common.py:48-92) with numpy / torch ops on the same decoded frames."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sequence(T, H, W, hwc=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (T, H, W, 3) if hwc else (T, 3, H, W), dtype=torch.uint8, generator=g)
    K = np.eye(4)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = 525.0 * W / 640, 525.0 * H / 480, 319.5 * W / 640, 239.5 * H / 480
    return frames, K


def _reference_sample(frames_hwc, K, idx3, H, W):
    """One sample as common.py builds it: ToTensor of the three frames, K / pinv(K) per scale in float64 -> float32."""
    out = {}
    for s in range(4):
        Ks = K.copy()
        Ks[0, :] *= (W // (2 ** s)) / W
        Ks[1, :] *= (H // (2 ** s)) / H
        out[("K", s)] = torch.from_numpy(Ks).float()
        out[("inv_K", s)] = torch.from_numpy(np.linalg.pinv(Ks)).float()
    to_tensor = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).float().div(255)
    for key, i in zip((("source_left", 0), ("target_image", 0), ("source_right", 0)), idx3):
        out[key] = to_tensor(frames_hwc[i])
    return out


@pytest.mark.parametrize("hwc", [True, False])
@pytest.mark.parametrize("out", ["float32", "uint8"])
def test_batcher_matches_dataset_getitem_and_collation(hwc, out):
    from dvsloss import GpuTripletBatcher
    T, H, W, B = 12, 48, 64, 5
    frames, K = _sequence(T, H, W, hwc)
    bat = GpuTripletBatcher(frames.cuda(), torch.from_numpy(K), out=out, generator=torch.Generator().manual_seed(3))
    assert len(bat) == T - 6                                          # max_size 3 on both sides, as the reference
    idx = bat.draw_indices(B)
    assert idx.shape == (B, 3) and int(idx.min()) >= 0 and int(idx.max()) < T
    d1, d2 = idx[:, 1] - idx[:, 0], idx[:, 2] - idx[:, 1]
    assert d1.min() >= 1 and d1.max() <= 3 and d2.min() >= 1 and d2.max() <= 3
    sample = bat.batch(idx)
    hwc_np = frames.numpy() if hwc else frames.permute(0, 2, 3, 1).contiguous().numpy()
    ref = [_reference_sample(hwc_np, K, idx[b].tolist(), H, W) for b in range(B)]
    for key in ref[0]:
        want = torch.stack([r[key] for r in ref])                     # default_collate
        got = sample[key].cpu()
        if key[0] in ("K", "inv_K"):
            assert got.dtype == torch.float32 and torch.allclose(got, want, rtol=1e-6, atol=1e-9), key
        elif out == "float32":
            assert got.dtype == torch.float32 and torch.equal(got, want), key       # x / 255 is exact
        else:
            assert got.dtype == torch.uint8 and torch.equal(got.float().div(255), want), key


def test_batcher_feeds_the_loss_and_uint8_path_agrees():
    """A batch straight from the batcher drives the fused loss; bytes and ToTensor'ed floats give identical losses."""
    from dvsloss import GpuTripletBatcher, view_synthesis_loss
    T, H, W, B = 10, 64, 96, 2
    frames, K = _sequence(T, H, W, True, seed=1)
    dev = torch.device("cuda:0")
    idx = torch.tensor([[0, 1, 3], [2, 5, 6]], dtype=torch.int32)
    outs = []
    for out in ("float32", "uint8"):
        s = GpuTripletBatcher(frames.cuda(), torch.from_numpy(K), out=out).batch(idx)
        g = torch.Generator(device=dev).manual_seed(0)
        disps = [torch.rand(B, 1, H >> k, W >> k, device=dev, generator=g) for k in range(4)]
        Ts = [torch.eye(4, device=dev).repeat(B, 1, 1) for _ in range(2)]
        Ts[0][:, 0, 3], Ts[1][:, 0, 3] = 0.01, -0.01
        outs.append(view_synthesis_loss(disps, s[("target_image", 0)], [s[("source_left", 0)], s[("source_right", 0)]],
                                        s[("K", 0)], s[("inv_K", 0)], Ts, noise=None))
    assert torch.isfinite(outs[0][0]) and torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_batcher_color_jitter_is_shared_by_the_three_frames():
    from dvsloss import GpuTripletBatcher
    T, H, W = 10, 32, 48
    frames = torch.full((T, H, W, 3), 128, dtype=torch.uint8)
    bat = GpuTripletBatcher(frames.cuda(), torch.eye(4), augment=True, generator=torch.Generator().manual_seed(0))
    s = bat.sample(8)
    l, t, r = s[("source_left", 0)], s[("target_image", 0)], s[("source_right", 0)]
    assert torch.equal(l, t) and torch.equal(t, r)                  # identical frames + identical jitter parameters
    assert float(t.min()) >= 0 and float(t.max()) <= 1
    changed = (t.flatten(1) - 128 / 255).abs().amax(1) > 1e-6
    assert 0 < int(changed.sum()) < 8                               # applied with probability 0.5 per sample
