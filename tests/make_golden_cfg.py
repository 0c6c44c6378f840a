"""Configurations of the committed golden frame dumps (shared by make_golden.py and the GPU test)."""
from eidola_b200 import abi, scenes

CONFIGS = {
    # name: (scene maker, (W, H), frames, RtxState overrides[, extras: sun_sky = SunAndSky keyword overrides])
    # every dump also holds `display`: post.frag (default Tonemapper) of the last frame
    "c1_cube": (scenes.cube_scene, (96, 96), 1, dict(ReSTIRState=abi.eRIS, RISSampleNum=1, denoise=0)),
    "c2_cornell": (scenes.cornell_scene, (128, 72), 3, dict(maxDepth=3)),
    "room": (scenes.small_room, (96, 64), 3, dict()),
    "cube_sunsky": (scenes.cube_scene, (96, 80), 2, dict(environmentProb=0.25, fireflyClampThreshold=50.0, maxDepth=3),
                    dict(sun_sky=dict(in_use=1, sun_direction=(0.3, 0.5, 0.4), haze=1.0))),
}


def sun_sky_of(entry):
    """SunAndSky of a CONFIGS entry (None when it has none)."""
    if len(entry) < 5 or "sun_sky" not in entry[4]:
        return None
    kw = dict(entry[4]["sun_sky"])
    if "sun_direction" in kw:
        kw["sun_direction"] = abi.Vec3(*kw["sun_direction"])
    return abi.default_sun_and_sky(**kw)
