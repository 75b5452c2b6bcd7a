"""Dataset schema mirrors the reference's features.py (keys, shapes, dtypes, dimension names)."""
from mujoco_manip_b200 import features as F


def test_schema_shapes_and_names():
    assert len(F.FEATURES) == 19
    for k, names in F.DIM_NAMES.items():
        assert k in F.FEATURES and F.FEATURES[k]["shape"] == (len(names),), k
        assert F.FEATURES[k]["dtype"] == "float32"
    assert F.FEATURES["observation.images.overhead"]["shape"] == (224, 224, 3)
    assert F.FEATURES["observation.phase_description"]["dtype"] == "string"
    assert F.DIM_NAMES["observation.state"] == ["ee_x", "ee_y", "ee_z", "gripper", "q0", "q1", "q2", "q3", "q4", "q5", "q6"]
    assert F.DIM_NAMES["observation.keypoints_wrist"][6:8] == ["bin_red_u", "bin_red_v"]
    assert F.DIM_NAMES["next.reward"] == ["total", "reach_obj", "pick_obj", "reach_target", "place_obj", "reach_home"]
    obs = sorted(b - a for a, b in F._OBS_FEATURES.values())
    assert sum(obs) == 85
    assert sum(b - a for a, b in F._ACT_FEATURES.values()) == 36


def test_phase_strings():
    assert F.phase_description(1, "obj_red", "bin_blue") == "idle"
    assert F.phase_description(6, "obj_red", "bin_blue") == "transporting the red cube to the blue bin"
    assert F.phase_description(10, "obj_red", "bin_blue") == "retreating to neutral position"


def test_xml_path_accepts_only_the_bundled_scene(tmp_path):
    """gym_env.py:64 / env.py:15-69: the reference's default xml_path must be accepted; other scenes are refused loudly."""
    import os

    import pytest

    from mujoco_manip_b200.constants import check_scene_xml

    check_scene_xml(None)
    ref = "/root/reference/mujoco_manip/data/pick_and_place_scene.xml"
    if os.path.exists(ref):  # only in the build container
        check_scene_xml(ref)
        other = tmp_path / "scene.xml"
        other.write_text(open(ref).read().replace('timestep="0.002"', 'timestep="0.004"'))
        with pytest.raises(ValueError):
            check_scene_xml(str(other))
    with pytest.raises(FileNotFoundError):
        check_scene_xml(str(tmp_path / "missing.xml"))
