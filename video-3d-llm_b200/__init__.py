"""video-3d-llm_b200 -- B200-native (sm_100a) position-aware video front end of Video-3D LLM.

Only the hot path named by BASELINE.json:north_star lives here:
  csrc/                hand-written CUDA kernels + the C ABI (include/v3d.h) -> libv3d.so
  cabi.py              ctypes binding of the C ABI (fails loudly when the library is missing)
  video_utils.py       mirror of llava/video_utils.py (unproject, VideoProcessor, merge_video_dict)
  position_encoding.py mirror of llava/model/position_encoding.py
  coords.py            mirror of the coordinate helpers of llava/model/llava_arch.py:213-272
  frontend.py          the fused depth -> pooled xyz -> PE (+add) entry point
  mc_sampling.py       mirror of scripts/3d/preprocessing/max_coverage_sampling.py
  sharding.py          scene sharding across ranks + all-gather of per-scene records
  launch.py            overlay launcher that lets the reference scripts run unchanged
  synthetic.py         deterministic ScanNet-shaped inputs for tests and bench.py
"""
__version__ = "0.1.0"
