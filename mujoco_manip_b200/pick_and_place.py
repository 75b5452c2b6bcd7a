"""Scripted pick-and-place FSM with the reference's class surface (mujoco_manip/pick_and_place.py).

The state machine itself runs on the GPU (`mm_fsm_plan`, csrc/mm_env.h fsm_plan_one); this class is the
host-side view for one env of a PickPlaceGymEnv: same State / Phase enums, `plan`, `update`,
`_actuate`, `state`, `phase`, `phase_description`, `target_pos`, `gripper_val`, `is_done`.
"""
from __future__ import annotations

from enum import Enum, auto

import numpy as np

from .constants import BINS, OBJECTS


class State(Enum):
    IDLE = auto()
    PRE_GRASP = auto()
    GRASP = auto()
    CLOSE_GRIPPER = auto()
    LIFT = auto()
    MOVE_TO_BIN = auto()
    SETTLE_AT_BIN = auto()
    LOWER_TO_BIN = auto()
    RELEASE = auto()
    RETREAT = auto()
    DONE = auto()


class Phase(Enum):
    IDLE = "idle"
    APPROACHING = "approaching"
    GRASPING = "grasping"
    LIFTING = "lifting"
    TRANSPORTING = "transporting"
    PLACING = "placing"
    RETREATING = "retreating"
    DONE = "done"


_STATE_TO_PHASE = {
    State.IDLE: Phase.IDLE, State.PRE_GRASP: Phase.APPROACHING, State.GRASP: Phase.GRASPING,
    State.CLOSE_GRIPPER: Phase.GRASPING, State.LIFT: Phase.LIFTING, State.MOVE_TO_BIN: Phase.TRANSPORTING,
    State.SETTLE_AT_BIN: Phase.TRANSPORTING, State.LOWER_TO_BIN: Phase.PLACING, State.RELEASE: Phase.PLACING,
    State.RETREAT: Phase.RETREATING, State.DONE: Phase.DONE,
}

TASKS = [("obj_red", "bin_red"), ("obj_green", "bin_green"), ("obj_blue", "bin_blue")]

# heights of the hand frame / timers (pick_and_place.py:62-74); the device FSM holds the same values
PRE_GRASP_HEIGHT, GRASP_HEIGHT, LIFT_HEIGHT = 0.44, 0.36, 0.55
TRANSIT_HEIGHT, RELEASE_HEIGHT, RETREAT_HEIGHT = 0.55, 0.45, 0.55
GRIPPER_SETTLE_STEPS, BIN_SETTLE_STEPS, TRANSIT_SPEED = 150, 100, 0.001


class PickAndPlaceTask:
    """FSM view bound to env 0 of the PickPlaceVecEnv behind a PickPlaceGymEnv.

    `tasks` lists the (object, bin) pairs to execute in order (default TASKS, pick_and_place.py:91); the list is the
    FSM's own (device field `fsm_tasks`) and does not touch the env's task (reward, success test, one-hot observations),
    exactly as in the reference.  All transitions, including IDLE -> PRE_GRASP of the next pair, happen on the device.
    """

    MAX_TASKS = 9

    def __init__(self, env, robot, controller, tasks=None):
        import torch

        self.env, self.robot, self.controller = env, robot, controller
        self._tasks = list(tasks or TASKS)
        if len(self._tasks) > self.MAX_TASKS:
            raise ValueError(f"at most {self.MAX_TASKS} tasks")
        self._vec = env._vec
        v = self._vec
        row = [len(self._tasks)]
        for o, b in self._tasks:
            row += [OBJECTS.index(o), BINS.index(b)]
        row += [0] * (20 - len(row))
        v.state["fsm_tasks"][0] = torch.tensor(row, dtype=torch.int32, device=v.device)
        v.state["fsm_i"][0] = torch.tensor([1, 0, 0, 1, 0], dtype=torch.int32, device=v.device)
        v.state["fsm_f"][0].zero_()

    def _fsm(self):
        return self._vec.state["fsm_i"][0].cpu().numpy()

    @property
    def state(self) -> State:
        return State(int(self._fsm()[0]))

    @property
    def task_index(self) -> int:
        return int(self._fsm()[1])

    @property
    def settle_counter(self) -> int:
        return int(self._fsm()[2])

    @property
    def is_done(self) -> bool:
        return self.state == State.DONE

    @property
    def phase(self) -> Phase:
        return _STATE_TO_PHASE[self.state]

    @property
    def target_pos(self):
        f = self._fsm()
        return self._vec.state["fsm_f"][0, :3].cpu().numpy().copy() if f[4] else None

    @property
    def _target_pos(self):
        return self.target_pos

    @property
    def _gripper_open(self) -> bool:
        return bool(self._fsm()[3])

    @property
    def gripper_val(self) -> float:
        return 1.0 if self._gripper_open else 0.0

    def _obj_name(self) -> str:
        return self._tasks[min(self.task_index, len(self._tasks) - 1)][0]

    def _bin_name(self) -> str:
        return self._tasks[min(self.task_index, len(self._tasks) - 1)][1]

    @property
    def phase_description(self) -> str:
        """pick_and_place.py:127-149"""
        ph = self.phase
        if ph in (Phase.IDLE, Phase.DONE):
            return "idle"
        if ph == Phase.RETREATING:
            return "retreating to neutral position"
        oc, bc = self._obj_name().replace("obj_", ""), self._bin_name().replace("bin_", "")
        return {Phase.APPROACHING: f"approaching the {oc} cube", Phase.GRASPING: f"grasping the {oc} cube",
                Phase.LIFTING: f"lifting the {oc} cube",
                Phase.TRANSPORTING: f"transporting the {oc} cube to the {bc} bin",
                Phase.PLACING: f"placing the {oc} cube in the {bc} bin"}.get(ph, "idle")

    def plan(self, n_steps: int = 1) -> str:
        """One FSM tick covering n_steps physics steps (at most one transition), pick_and_place.py:167-277.
        Returns the reference's status string, derived from the transition the device FSM made."""
        s0, ti = self.state, self.task_index
        self._vec.fsm_plan(n_steps)
        s1, cnt = self.state, self.settle_counter
        if s0 == State.DONE or (s0 == State.IDLE and s1 == State.DONE):
            return "All objects placed!"
        if s0 == State.RETREAT:
            return "Ready for next object" if s1 == State.IDLE else "Retreating"
        if s0 == State.RELEASE:
            return "Retreating to neutral position" if s1 == State.RETREAT else f"Releasing ({cnt})"
        obj, bn = self._tasks[min(ti, len(self._tasks) - 1)]
        moved = s1 != s0
        return {
            State.IDLE: (f"Moving to pre-grasp above {obj}",) * 2,
            State.PRE_GRASP: (f"Descending to grasp {obj}", f"Approaching pre-grasp for {obj}"),
            State.GRASP: (f"Closing gripper on {obj}", f"Descending to {obj}"),
            State.CLOSE_GRIPPER: (f"Lifting {obj}", f"Gripping {obj} ({cnt})"),
            State.LIFT: (f"Moving {obj} to {bn}", f"Lifting {obj}"),
            State.MOVE_TO_BIN: (f"Settling above {bn}", f"Transporting {obj}"),
            State.SETTLE_AT_BIN: (f"Lowering {obj} into {bn}", f"Settling above {bn} ({cnt})"),
            State.LOWER_TO_BIN: (f"Releasing {obj} into {bn}", f"Lowering to {bn}"),
        }[s0][0 if moved else 1]

    def _actuate(self) -> None:
        if self._gripper_open:
            self.robot.open_gripper()
        else:
            self.robot.close_gripper()
        t = self.target_pos
        if t is not None:
            self.robot.set_arm_ctrl(self.controller.compute(t))

    def update(self) -> str:
        status = self.plan(1)
        self._actuate()
        return status
