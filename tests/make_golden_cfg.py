"""Configurations of the committed golden frame dumps (shared by make_golden.py and the GPU test)."""
from eidola_b200 import abi, scenes

CONFIGS = {
    # name: (scene maker, (W, H), frames, RtxState overrides)
    "c1_cube": (scenes.cube_scene, (96, 96), 1, dict(ReSTIRState=abi.eRIS, RISSampleNum=1, denoise=0)),
    "c2_cornell": (scenes.cornell_scene, (128, 72), 3, dict(maxDepth=3)),
    "room": (scenes.small_room, (96, 64), 3, dict()),
}
