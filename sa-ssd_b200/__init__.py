"""sassd_b200 — B200-native implementation of the SA-SSD point-cloud inference
hot path (voxelize -> sparse 3D conv backbone -> BEV neck -> SSD rotate head ->
PSWarp rescoring -> rotated NMS) behind the reference's mmdet.models operator
API.  Hand-written sm_100a CUDA lives in ``csrc/`` behind the C ABI declared in
``include/sassd_b200.h``; this package is the host-side mirror of the reference
interface.  Import name: ``sassd_b200`` (alias of the ``sa-ssd_b200/`` directory).
"""
__version__ = "0.1.0"
