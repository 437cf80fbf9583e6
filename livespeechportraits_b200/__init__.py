"""B200-native implementation of LiveSpeechPortraits' per-frame renderer (``Feature2Face_G``) and the steps either
side of it.  Public names:

* ``Feature2Face_G`` / ``install`` - drop-in for ``models/feature2face_G.py`` (generator.py)
* ``ClipRenderer`` - batched host-to-host render loop (pipeline.py)
* ``ShardedRenderer`` / ``partition`` - one process per GPU, frames partitioned, one gather (parallel.py)
* ``render_to_video`` - landmark tracks -> video file, the demo.py loop body without JPEGs (video.py)

Importing the package does not touch CUDA; the native library is loaded (and built if missing) on first use.
"""
from .generator import Feature2Face_G, install  # noqa: F401
from .parallel import ShardedRenderer, partition  # noqa: F401
from .pipeline import ClipRenderer  # noqa: F401
from .video import render_to_video  # noqa: F401

__all__ = ["Feature2Face_G", "install", "ClipRenderer", "ShardedRenderer", "partition", "render_to_video"]
