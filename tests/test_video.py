"""Row N3: the video writer that stands where demo.py:35-45 / util/visualizer.py:120-143 stand.
CPU part: the writer thread produces a file cv2 reads back with the right frame count, size and channel order.
GPU part: landmark tracks -> video through the generator's kernels."""
import os

import numpy as np
import pytest
import torch


def _read_all(path):
    import cv2
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    fps = cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    return frames, fps


def test_writer_thread_writes_bgr_frames_in_order(tmp_path):
    from livespeechportraits_b200 import video as V
    path = str(tmp_path / "t.avi")
    wt = V._WriterThread(V._open_writer(path, 60, (64, 48), "DIVX"))      # demo.py:36: 60 fps DIVX
    rgb = np.zeros((10, 48, 64, 3), dtype=np.uint8)
    for i in range(10):
        rgb[i, :, :, 0] = 255 if i % 2 == 0 else 0            # even frames red, odd frames blue (RGB order in)
        rgb[i, :, :, 2] = 0 if i % 2 == 0 else 255
    wt.put(rgb[:4])
    wt.put(rgb[4:])
    wt.close()
    frames, fps = _read_all(path)
    assert len(frames) == 10 and frames[0].shape == (48, 64, 3) and abs(fps - 60) < 1e-3
    for i, f in enumerate(frames):                              # cv2 reads BGR
        b, g, r = [float(f[..., c].mean()) for c in range(3)]
        assert (r > 180 and b < 80) if i % 2 == 0 else (b > 180 and r < 80), (i, b, g, r)
    assert wt.frames == 10
    assert V.mux_audio(path, path, str(tmp_path / "x.avi")) in (False, True)     # False here: no ffmpeg binary in the image


@pytest.mark.gpu
def test_render_to_video_from_landmark_tracks(tmp_path):
    import types
    from livespeechportraits_b200 import Feature2Face_G, render_to_video
    from oracle import f2f_oracle as O
    from oracle import raster_oracle as RO
    net = Feature2Face_G(types.SimpleNamespace(isTrain=False, size="normal", n_downsample_G=8, ngf=64, fp16=0), precision="parity")
    net.load_state_dict(O.make_state_dict("normal", "B"))
    net = net.cuda().eval()
    n, size = 11, (256, 256)
    lm, sh = RO.make_landmarks(n, size, seed=5)
    _, cand = O.make_inputs(1, 256, 256, seed=6)
    cand_d = cand[:1].cuda()
    path, fpath = str(tmp_path / "clip.avi"), str(tmp_path / "maps.avi")
    info = render_to_video(net, lm, sh, cand_d, path, size=size, fps=60, batch=4, feature_maps_path=fpath)
    assert info["frames"] == n and info["audio_muxed"] is False and os.path.getsize(path) > 0
    frames, fps = _read_all(path)
    assert len(frames) == n and frames[0].shape == (256, 256, 3) and abs(fps - 60) < 1e-3
    # the frames in the file are the generator's uint8 images (lossy codec: compare loosely, in BGR order)
    fm = net.draw_feature_maps(torch.from_numpy(lm).cuda(), torch.from_numpy(sh).cuda(), size)
    ref = net.render_image(fm, cand_d).cpu().numpy()
    for i in (0, n - 1):
        d = np.abs(frames[i][..., ::-1].astype(np.int16) - ref[i].astype(np.int16))
        assert d.mean() < 12, d.mean()
    maps, _ = _read_all(fpath)
    assert len(maps) == n
    m0 = (fm[0, 0].cpu().numpy() * 255).astype(np.int16)
    assert np.abs(maps[0][..., 0].astype(np.int16) - m0).mean() < 12
