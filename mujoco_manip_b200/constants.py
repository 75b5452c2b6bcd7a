"""Task vocabulary and timing constants of the pick-and-place scene.

Same names and values as the reference's mujoco_manip/constants.py:3-37 so that code written against
the reference keeps working; indices into OBJECTS / BINS are what the CUDA kernels receive as
(object index, bin index).
"""
OBJECTS = ["obj_red", "obj_green", "obj_blue"]
BINS = ["bin_red", "bin_green", "bin_blue"]


def _colour(body: str) -> str:
    return body.split("_", 1)[1]


ALL_TASKS = [(obj, b) for obj in OBJECTS for b in BINS]
MATCH_TASKS = [t for t in ALL_TASKS if _colour(t[0]) == _colour(t[1])]
CROSS_TASKS = [t for t in ALL_TASKS if _colour(t[0]) != _colour(t[1])]
TASK_SETS = {"all": ALL_TASKS, "match": MATCH_TASKS, "cross": CROSS_TASKS}

IMAGE_SIZE = 224
CONTROL_FPS = 30
PHYSICS_DT = 0.002
ACTION_REPEAT = 16  # physics substeps per control step (about 31 Hz)
MAX_EPISODE_STEPS = 500

KEYPOINT_BODIES = OBJECTS + BINS + ["hand"]

# spawn defaults of randomization.py:15-18 / gym_env.py:73-74
SPAWN_X_RANGE = (-0.20, 0.20)
SPAWN_Y_RANGE = (0.30, 0.45)
OBJ_SPAWN_Z = 0.26
MIN_OBJ_SEPARATION = 0.08
MAX_REJECTION_ATTEMPTS = 1000


# The CUDA kernels are specialised (tools/modelc.py) to the reference's bundled scene; `xml_path` arguments are checked
# against the SHA-256 of the files the tables were generated from (mujoco_manip/data/pick_and_place_scene.xml, which
# includes franka_emika_panda/panda.xml)
SCENE_XML_SHA256 = "14ee126fef8d1b782ddea60b1e94505291eb3ff8ed673774a961fd80417703f8"
PANDA_XML_SHA256 = "96ad67da03710f17f798c9478fd9e9efdf24a3bf8359f05e456dd9fb158ea273"


def check_scene_xml(xml_path) -> None:
    """Accept `xml_path` (gym_env.py:64, env.py:15-69) when it is the bundled scene the kernels were compiled from.

    None is the default scene.  A path is read as the reference reads it (FileNotFoundError when missing) and must have
    the bundled scene's content; any other scene raises ValueError because its model tables do not exist on the device.
    """
    if xml_path is None:
        return
    import hashlib

    with open(xml_path, "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()
    if digest != SCENE_XML_SHA256:
        raise ValueError(f"{xml_path}: not the bundled pick_and_place_scene.xml (sha256 {digest[:12]}...); the CUDA kernels "
                         "are compiled for that scene only (tools/modelc.py regenerates the tables for another one)")


def task_indices(task) -> tuple[int, int]:
    """(object index, bin index) of an (obj_name, bin_name) pair."""
    obj, b = task
    return OBJECTS.index(obj), BINS.index(b)
