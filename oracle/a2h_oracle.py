"""CPU oracle for row N4: the Audio2Headpose autoregressive loop.  TEST INFRASTRUCTURE ONLY - nothing
under ``livespeechportraits_b200/`` may import this module (same rule as f2f_oracle.py).

Functional restatement, on torch ATen CPU ops, of

* ``Audio2Headpose.forward``                     models/audio2headpose.py:41-53  (audio_downsample :17-22)
* ``WaveNet.forward`` / ``residual_block.forward``  models/networks.py:199-227, 303-326 (constructor :103-191, :260-301)
* ``Sample_GMM``                                  models/losses.py:68-112
* ``Audio2HeadposeModel.generate_sequences``      models/audio2headpose_model.py:133-187 (WaveNet decoder, fill_zero=True)

with the options of options/base_options_audio2headpose.py:65-78 and the test-time default ``time_frame_length = 1``
(options/test_audio2headpose_options.py:17).  The restatement follows the REFERENCE algorithm - a full 255-step WaveNet
forward per generated frame - not the incremental form the CUDA kernel uses; tests/test_a2h_oracle.py pins it against the
live reference modules (bit-exact on this machine) and against golden vectors generated from them
(oracle/make_golden_a2h.py).

The only randomness of the loop is inside ``Sample_GMM`` (``torch.multinomial`` then ``torch.randn`` per frame, global
CPU generator).  ``reference_noise`` reproduces those draws, so the loop is a deterministic function of
``(weights, audio features, noise)`` and can be compared across implementations.
"""
from __future__ import annotations

import os
import sys
import types
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("LSP_REFERENCE_ROOT", "/root/reference")


def default_opt(**over):
    """The fields of HeadposeOptions().parse() that this path reads (options/base_options_audio2headpose.py:38-78,
    options/test_audio2headpose_options.py:17), plus ``A2H_receptive_field`` as demo.py:163-166 sets it."""
    d = dict(
        model="audio2headpose", feature_decoder="WaveNet", loss="GMM", isTrain=False, gpu_ids=[], task="Audio2Headpose",
        APC_hidden_size=512, frame_future=15, time_frame_length=1,
        A2H_wavenet_residual_layers=7, A2H_wavenet_residual_blocks=2, A2H_wavenet_dilation_channels=128,
        A2H_wavenet_residual_channels=128, A2H_wavenet_skip_channels=256, A2H_wavenet_kernel_size=2,
        A2H_wavenet_use_bias=True, A2H_wavenet_cond=True, A2H_wavenet_cond_channels=512, A2H_wavenet_input_channels=12,
        A2H_GMM_ncenter=1, A2H_GMM_ndim=12, A2H_GMM_sigma_min=0.03,
        checkpoints_dir="/tmp", name="Audio2Headpose", load_epoch="none", verbose=False, smooth_loss=0, continue_train=False,
    )
    d.update(over)
    o = types.SimpleNamespace(**d)
    o.A2H_receptive_field = receptive_field(o)
    return o


def dilations(opt) -> Tuple[int, ...]:
    """networks.py:160-176: per block the dilation restarts at 1 and doubles per layer."""
    return tuple(2 ** i for _ in range(opt.A2H_wavenet_residual_blocks) for i in range(opt.A2H_wavenet_residual_layers))


def receptive_field(opt) -> int:
    """networks.py:147,175-176: 1 + sum over layers of (kernel_size - 1) * dilation."""
    return 1 + sum((opt.A2H_wavenet_kernel_size - 1) * d for d in dilations(opt))


def output_size(opt) -> int:
    return (2 * opt.A2H_GMM_ndim + 1) * opt.A2H_GMM_ncenter if opt.loss == "GMM" else opt.A2H_GMM_ndim      # audio2headpose.py:11-14


def state_dict_spec(opt) -> "OrderedDict[str, Tuple[int, ...]]":
    """Key grammar and shapes of ``Audio2Headpose(opt).state_dict()`` (pinned against the live module in the tests)."""
    H = opt.APC_hidden_size
    R, D, S = opt.A2H_wavenet_residual_channels, opt.A2H_wavenet_dilation_channels, opt.A2H_wavenet_skip_channels
    C, I, ks, O = opt.A2H_wavenet_cond_channels, opt.A2H_wavenet_input_channels, opt.A2H_wavenet_kernel_size, output_size(opt)
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    spec["audio_downsample.0.weight"] = (H, 2 * H)
    spec["audio_downsample.0.bias"] = (H,)
    for k, shp in (("weight", (H,)), ("bias", (H,)), ("running_mean", (H,)), ("running_var", (H,)), ("num_batches_tracked", ())):
        spec["audio_downsample.1." + k] = shp
    spec["audio_downsample.3.weight"] = (H, H)
    spec["audio_downsample.3.bias"] = (H,)
    spec["WaveNet.start_conv1.weight"] = (R, I, 1)
    spec["WaveNet.start_conv1.bias"] = (R,)
    spec["WaveNet.start_conv2.weight"] = (R, R, 1)
    spec["WaveNet.start_conv2.bias"] = (R,)
    for i in range(len(dilations(opt))):
        p = f"WaveNet.residual_blocks.{i}."
        for name, shp in (("filter_conv", (D, R, ks)), ("gate_conv", (D, R, ks)), ("residual_conv", (R, D, 1)),
                          ("skip_conv", (S, D, 1))):
            spec[p + name + ".weight"] = shp
            if opt.A2H_wavenet_use_bias:
                spec[p + name + ".bias"] = (shp[0],)
        for name in ("cond_filter_conv", "cond_gate_conv"):
            spec[p + name + ".weight"] = (D, C, 1)
            spec[p + name + ".bias"] = (D,)
    spec["WaveNet.end_conv_1.weight"] = (O, S, 1)
    spec["WaveNet.end_conv_1.bias"] = (O,)
    spec["WaveNet.end_conv_2.weight"] = (O, O, 1)
    spec["WaveNet.end_conv_2.bias"] = (O,)
    return spec


def make_state_dict(opt, recipe: str = "B", seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic weights (no checkpoint ships).  Recipe A = what ``networks.init_weights('normal', 0.02)`` leaves
    (networks.py:347-378: Conv/Linear weights N(0, 0.02), biases 0; BatchNorm1d untouched).  Recipe B = weights
    N(0, 1/fan_in) (unit-gain layers, so the feedback through the history matters), biases N(0, 0.1), randomised
    BatchNorm statistics - every term of the arithmetic is exercised.  numpy PCG64 in key order, so any machine
    regenerates identical bits."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for k, shp in state_dict_spec(opt).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0, dtype=torch.long)
            continue
        is_bn = k.startswith("audio_downsample.1.")
        if recipe == "A":
            if is_bn:
                v = np.ones(shp, np.float32) if k.endswith(("weight", "running_var")) else np.zeros(shp, np.float32)
            elif k.endswith("weight"):
                v = rng.normal(0.0, 0.02, shp).astype(np.float32)
            else:
                v = np.zeros(shp, np.float32)
        else:
            if is_bn:
                if k.endswith("weight"):
                    v = rng.normal(1.0, 0.1, shp)
                elif k.endswith("running_var"):
                    v = rng.uniform(0.5, 1.5, shp)
                else:
                    v = rng.normal(0.0, 0.2, shp)
                v = v.astype(np.float32)
            elif k.endswith("weight"):
                fan_in = int(np.prod(shp[1:]))
                v = rng.normal(0.0, 1.0 / np.sqrt(fan_in), shp).astype(np.float32)
            else:
                v = rng.normal(0.0, 0.1, shp).astype(np.float32)
        sd[k] = torch.from_numpy(np.ascontiguousarray(v))
    return sd


def make_audio_feats(n_frames: int, opt, seed: int = 1) -> np.ndarray:
    """Synthetic stand-in for the APC features demo.py:196-206 feeds (``[n, 2 * APC_hidden]`` after the reshape at
    audio2headpose_model.py:148): smooth in time like speech features, O(1) magnitude."""
    rng = np.random.Generator(np.random.PCG64(9000 + seed))
    base = rng.normal(0.0, 1.0, (n_frames + 8, 2 * opt.APC_hidden_size)).astype(np.float32)
    k = np.array([1, 2, 3, 2, 1], np.float32)
    k /= k.sum()
    sm = sum(k[j] * base[j:j + n_frames] for j in range(5))
    return np.ascontiguousarray(sm * 1.5, dtype=np.float32)


def reference_noise(n_frames: int, ndim: int, ncenter: int = 1, seed: int = 0) -> np.ndarray:
    """The ``torch.randn`` draws ``Sample_GMM`` makes for a clip when the global CPU generator is seeded with ``seed`` right
    before ``generate_sequences``: per frame one ``torch.multinomial`` (losses.py:87) then ``torch.randn(1, ndim)`` (:96)."""
    torch.manual_seed(seed)
    out = np.zeros((n_frames, ndim), np.float32)
    prob = torch.full((1, ncenter), 1.0 / ncenter)
    for i in range(n_frames):
        torch.multinomial(prob, num_samples=1, replacement=True)
        out[i] = torch.randn(1, ndim).float().numpy()[0]
    return out


# ----------------------------------------------------------------------------------------------
# The restatement
# ----------------------------------------------------------------------------------------------

def audio_downsample(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """audio2headpose.py:17-22 in eval mode: Linear(1024,512) - BatchNorm1d(512) - LeakyReLU(0.2) - Linear(512,512).  x: [n, 1024]."""
    y = F.linear(x, sd["audio_downsample.0.weight"], sd["audio_downsample.0.bias"])
    y = F.batch_norm(y, sd["audio_downsample.1.running_mean"], sd["audio_downsample.1.running_var"], sd["audio_downsample.1.weight"],
                     sd["audio_downsample.1.bias"], False, 0.1, 1e-5)
    y = F.leaky_relu(y, 0.2)
    return F.linear(y, sd["audio_downsample.3.weight"], sd["audio_downsample.3.bias"])


def wavenet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cond: torch.Tensor, opt) -> torch.Tensor:
    """networks.py:199-227 in eval mode (Dropout2d is the identity).  x: [b, ndim_in, T], cond: [b, cond_ch, T] -> [b, out_len, out_ch]."""
    act = lambda t: F.leaky_relu(t, 0.2)          # noqa: E731  networks.py:140-141 (activation='leakyrelu')
    b = lambda k: sd.get(k)                        # noqa: E731  biases absent when use_bias is False
    x = act(F.conv1d(x, sd["WaveNet.start_conv1.weight"], sd["WaveNet.start_conv1.bias"]))
    x = act(F.conv1d(x, sd["WaveNet.start_conv2.weight"], sd["WaveNet.start_conv2.bias"]))
    skip = 0
    ks = opt.A2H_wavenet_kernel_size
    for i, d in enumerate(dilations(opt)):
        p = f"WaveNet.residual_blocks.{i}."
        xp = F.pad(x, ((ks - 1) * d, 0))                                                   # networks.py:271,307
        filt = F.conv1d(xp, sd[p + "filter_conv.weight"], b(p + "filter_conv.bias"), dilation=d)
        gate = F.conv1d(xp, sd[p + "gate_conv.weight"], b(p + "gate_conv.bias"), dilation=d)
        filt = filt + F.conv1d(cond, sd[p + "cond_filter_conv.weight"], sd[p + "cond_filter_conv.bias"])
        gate = gate + F.conv1d(cond, sd[p + "cond_gate_conv.weight"], sd[p + "cond_gate_conv.bias"])
        z = torch.tanh(filt) * torch.sigmoid(gate)
        skip = skip + F.conv1d(z, sd[p + "skip_conv.weight"], b(p + "skip_conv.bias"))
        x = F.conv1d(z, sd[p + "residual_conv.weight"], b(p + "residual_conv.bias")) + x
    res = F.conv1d(act(skip), sd["WaveNet.end_conv_1.weight"], sd["WaveNet.end_conv_1.bias"])
    res = F.conv1d(act(res), sd["WaveNet.end_conv_2.weight"], sd["WaveNet.end_conv_2.bias"])
    res = res[:, :, -opt.time_frame_length:]
    return res.transpose(1, 2)


def audio2headpose_forward(sd, history: torch.Tensor, audio: torch.Tensor, opt) -> torch.Tensor:
    """audio2headpose.py:41-53.  history [b, T, ndim_in], audio [b, T, 2*APC_hidden] -> [b, out_len, out_ch]."""
    bs, item_len, nd = audio.shape
    down = audio_downsample(sd, audio.reshape(-1, nd)).reshape(bs, item_len, -1)
    return wavenet_forward(sd, history.permute(0, 2, 1), down.transpose(1, 2), opt)


def sample_gmm(params: torch.Tensor, noise: torch.Tensor, ncenter: int, ndim: int, sigma_scale: float,
               selected: Optional[torch.Tensor] = None) -> torch.Tensor:
    """losses.py:68-112 with the random draws passed in: ``noise`` [b*T, ndim] stands for ``torch.randn`` (:96) and
    ``selected`` [b*T] for the component ``torch.multinomial`` picked (:87; forced to 0 when ncenter == 1)."""
    bsz, T, _ = params.shape
    p = params.reshape(-1, (2 * ndim + 1) * ncenter)
    if selected is None:
        if ncenter != 1:
            raise ValueError("pass the selected component indices when ncenter > 1")
        selected = torch.zeros(bsz * T, dtype=torch.long)
    mu = p[:, ncenter:ncenter + ncenter * ndim]
    sigma = torch.exp(-p[:, ncenter + ncenter * ndim:]) * sigma_scale
    idx = (selected.view(-1, 1) * ndim + torch.arange(ndim).view(1, -1))
    sel_mu = torch.gather(mu, 1, idx)
    sel_sigma = torch.gather(sigma, 1, idx)
    return (noise * sel_sigma + sel_mu).reshape(bsz, T, -1)


def generate_sequences(sd, audio_feats: np.ndarray, pre_headpose: np.ndarray, noise: np.ndarray, opt,
                       sigma_scale: float = 0.3, return_params: bool = False):
    """audio2headpose_model.py:133-187 (WaveNet decoder, ``fill_zero=True``), with Sample_GMM's draws passed as ``noise``
    [nframe, ndim].  Returns ``pred_headpose`` [nframe, ndim] float64 like the reference (np.zeros default dtype) and,
    on request, the GMM parameters [nframe, out_ch] the network produced at every step."""
    ff, rf = opt.frame_future, opt.A2H_receptive_field
    ndim, ncenter = opt.A2H_GMM_ndim, opt.A2H_GMM_ncenter
    audio_feats = np.asarray(audio_feats, np.float32).reshape(-1, 2 * opt.APC_hidden_size)
    nframe = audio_feats.shape[0] - ff
    pred = np.zeros([nframe, ndim])
    params_out = np.zeros([nframe, output_size(opt)], np.float32)
    insert = np.repeat(audio_feats[0], rf - 1).reshape(-1, rf - 1).T                       # :153-155
    feats = np.concatenate([insert, audio_feats])
    hist = np.repeat(np.asarray(pre_headpose, np.float32), rf).reshape(-1, rf).T           # :157-159
    hist = torch.from_numpy(np.ascontiguousarray(hist)).unsqueeze(0).float()
    with torch.no_grad():
        for i in range(nframe):
            a = torch.from_numpy(feats[i + ff:i + ff + rf]).unsqueeze(0).float()            # :172-173
            preds = audio2headpose_forward(sd, hist, a, opt)                                 # :176
            if opt.loss == "GMM":
                data = sample_gmm(preds, torch.from_numpy(noise[i:i + 1]).float(), ncenter, ndim, sigma_scale)
            else:
                data = preds
            params_out[i] = preds[0, 0].numpy()
            pred[i] = data[0, 0].numpy()                                                     # :186
            hist = torch.cat((hist[:, 1:, :], data), dim=1)                                  # :187
    return (pred, params_out) if return_params else pred


# ----------------------------------------------------------------------------------------------
# The live reference (this container only; /root/reference does not exist on the GPU box)
# ----------------------------------------------------------------------------------------------

def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "audio2headpose_model.py"))


def reference_model(opt, sd: Optional[Dict[str, torch.Tensor]] = None):
    """The UNMODIFIED ``Audio2HeadposeModel`` (models/audio2headpose_model.py) on the CPU, eval mode, optionally loaded
    with ``sd`` (strict)."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from models.audio2headpose_model import Audio2HeadposeModel  # type: ignore
        m = Audio2HeadposeModel(opt)
    if sd is not None:
        m.Audio2Headpose.load_state_dict(sd, strict=True)
    m.eval()
    return m
