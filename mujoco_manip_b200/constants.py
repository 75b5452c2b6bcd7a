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


def task_indices(task) -> tuple[int, int]:
    """(object index, bin index) of an (obj_name, bin_name) pair."""
    obj, b = task
    return OBJECTS.index(obj), BINS.index(b)
