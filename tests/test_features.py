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


def test_dataset_row_assembly_and_parquet_round_trip(tmp_path):
    """Host side of the dataset writer (no GPU): rows are cut where the FSM was still running before plan(), the
    feature slices follow FEATURES, phase strings come from the FSM state after plan(), and a shard reads back equal."""
    import numpy as np

    from mujoco_manip_b200 import dataset
    from mujoco_manip_b200.features import FEATURES

    T = 7
    rng = np.random.default_rng(0)
    obs = rng.normal(size=(T, 85)).astype(np.float32)
    enc = rng.normal(size=(T, 36)).astype(np.float32)
    rc = rng.uniform(size=(T, 6)).astype(np.float32)
    state = np.array([2, 2, 3, 4, 10, 11, 11])
    running = np.array([1, 1, 1, 1, 1, 1, 0], dtype=bool)  # the frame whose plan() reached DONE is still recorded
    rows = dataset.assemble_episode_rows(obs, enc, state, running, rc, ("obj_green", "bin_red"), set(dataset.STATE_FEATURES))
    assert len(rows["frame_index"]) == 6 and rows["task"][0] == "Pick green object and place in red bin"
    for k in dataset.STATE_FEATURES:
        if FEATURES[k]["dtype"] == "float32":
            assert rows[k].shape == (6,) + tuple(FEATURES[k]["shape"]) and rows[k].dtype == np.float32, k
    assert np.array_equal(rows["observation.state"], obs[:6, :11]) and np.array_equal(rows["action.ee.pos_rot6d_g_rel"], enc[:6, 26:36])
    assert rows["observation.phase_description"] == ["approaching the green cube"] * 2 + ["grasping the green cube"] * 2 + \
        ["retreating to neutral position", "idle"]
    rows["episode_index"] = np.full(6, 3, dtype=np.int64)
    (tmp_path / "data").mkdir()
    dataset.write_parquet(str(tmp_path / "data" / "chunk-00000.parquet"), rows)
    back = dataset.read_episode(str(tmp_path), 3)
    assert np.array_equal(back["next.reward"], rc[:6]) and back["observation.phase_description"] == rows["observation.phase_description"]
    assert dataset.episode_seeds(42, 4) == [2684470948, 4091952314, 233227757, 3276785861]  # SURVEY 3.4 [DERIVED]
