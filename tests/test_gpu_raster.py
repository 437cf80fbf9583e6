"""GPU parity of the feature-map rasteriser through the C ABI (lspg_draw_feature_maps):
bit-exact against the oracle, against cv2 (the reference's dependency) and against the golden maps drawn by the
reference's own methods."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from oracle import f2f_oracle as O
from oracle import raster_oracle as R

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_NET = {}


def get_net():
    from livespeechportraits_b200.generator import Feature2Face_G
    if "n" not in _NET:
        opt = types.SimpleNamespace(isTrain=False, size="normal", n_downsample_G=8, ngf=64, fp16=0)
        net = Feature2Face_G(opt, precision="parity")
        net.load_state_dict(O.make_state_dict("normal", "B"))
        _NET["n"] = net.cuda().eval()
    return _NET["n"]


def gpu_maps(lm, sh, size):
    net = get_net()
    out = net.draw_feature_maps(torch.from_numpy(lm).cuda(), None if sh is None else torch.from_numpy(sh).cuda(), size)
    assert out.shape == (lm.shape[0], 1, size[1], size[0]) and out.dtype == torch.float32
    return out.cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "raster_*.npz"))), ids=os.path.basename)
def test_golden_maps(path):
    g = np.load(path)
    w, h = (int(v) for v in g["size"])
    lm, sh = g["landmarks"], (g["shoulders"] if g["shoulders"].size else None)
    got = gpu_maps(lm, sh, (w, h))
    assert set(np.unique(got)) <= {0.0, 1.0}
    for b in range(lm.shape[0]):
        ref = np.unpackbits(g["packed"][b])[: w * h].reshape(h, w).astype(np.float32)
        assert np.array_equal(got[b, 0], ref), (os.path.basename(path), b)


def test_random_segments_including_clipped_and_degenerate_ones():
    """Every landmark is an independent random point, so the 72 + 16 segments of a frame are random segments: long, short,
    zero-length (the part list repeats landmarks 18 and 24), far outside the image, negative fractional coordinates."""
    rng = np.random.default_rng(0)
    for w, h, margin in [(64, 64, 0), (64, 64, 40), (96, 40, 30), (512, 512, 300), (256, 512, 5)]:
        lm = rng.uniform(-margin, max(w, h) + margin, (12, 73, 2)).astype(np.float32)
        sh = rng.uniform(-margin, max(w, h) + margin, (12, 18, 2)).astype(np.float32)
        lm[0, :10] = rng.uniform(-0.99, 0.99, (10, 2))                 # int() truncates toward zero: all become (0, 0)
        got = gpu_maps(lm, sh, (w, h))
        for b in range(lm.shape[0]):
            ref = R.draw_feature_map_cv2(lm[b], (w, h), sh[b])
            assert np.array_equal(got[b, 0], (ref > 0).astype(np.float32)), (w, h, margin, b)
        b = 3
        assert np.array_equal(got[b, 0], (R.draw_feature_map(lm[b], (w, h), sh[b]) > 0).astype(np.float32))


def test_shoulder_variants_and_reuse_of_the_output_buffer():
    lm, sh = R.make_landmarks(2, (256, 256), seed=5, spill=0.1)
    a = gpu_maps(lm, None, (256, 256))
    b6 = gpu_maps(lm, sh[:, :6].copy(), (256, 256))
    for i in range(2):
        assert np.array_equal(a[i, 0] * 255, R.draw_feature_map(lm[i], (256, 256), None))
        assert np.array_equal(b6[i, 0] * 255, R.draw_feature_map(lm[i], (256, 256), sh[i, :6]))
    net = get_net()
    out = torch.full((2, 1, 256, 256), 7.0, device="cuda")           # stale contents must be overwritten
    net.draw_feature_maps(torch.from_numpy(lm).cuda(), None, (256, 256), out=out)
    assert np.array_equal(out.cpu().numpy(), a)
    with pytest.raises(ValueError):
        net.draw_feature_maps(torch.zeros(2, 70, 2, device="cuda"), None, (256, 256))


def test_rasterised_maps_drive_the_generator():
    """demo.py:262-266 with both steps on the GPU: landmarks -> maps -> frames equals maps drawn on the host -> frames."""
    net = get_net()
    lm, sh = R.make_landmarks(2, (256, 256), seed=9)
    _, cand = O.make_inputs(1, 256, 256, seed=4)
    fm_dev = net.draw_feature_maps(torch.from_numpy(lm).cuda(), torch.from_numpy(sh).cuda(), (256, 256))
    fm_host = torch.from_numpy(np.stack([R.feature_map_tensor(lm[b], (256, 256), sh[b]) for b in range(2)]))
    assert torch.equal(fm_dev.cpu(), fm_host)
    a = net.render(fm_dev, cand.cuda())
    b = net.render(fm_host.cuda(), cand.cuda())
    assert torch.equal(a, b)


def test_clip_renderer_from_landmarks_matches_host_drawn_maps():
    from livespeechportraits_b200.pipeline import ClipRenderer
    net = get_net()
    lm, sh = R.make_landmarks(5, (256, 256), seed=10, spill=0.05)
    _, cand = O.make_inputs(1, 256, 256, seed=4)
    fm_host = torch.from_numpy(np.stack([R.feature_map_tensor(lm[b], (256, 256), sh[b]) for b in range(5)])).pin_memory()
    ref = torch.empty((5, 256, 256, 3), dtype=torch.uint8).pin_memory()
    got = torch.empty((5, 256, 256, 3), dtype=torch.uint8).pin_memory()
    r = ClipRenderer(net, batch=2, uint8=True)
    r.render_clip(fm_host, cand.cuda(), ref)                                  # ragged last batch of 1
    r.render_clip_from_landmarks(torch.from_numpy(lm).pin_memory(), torch.from_numpy(sh).pin_memory(), cand.cuda(), got, (256, 256))
    assert torch.equal(ref, got)
