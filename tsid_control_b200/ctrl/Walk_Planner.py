"""WalkPlanner — what ref:ctrl/Walk_Planner.py:14-32 sets out to do.  The reference class is an unfinished stub
(`len(footsteps - 2)` at :23 raises, `step_duration` at :26 is undefined, nothing is returned); this version
keeps its plan(): for every i, a swing trajectory from footsteps[i] to footsteps[i + 2] (the same foot's next
placement) over one step_duration, and returns them with the support schedule that update_tasks
(ref:ctrl/WalkController.py:189-206) consumes.
"""
from typing import List

from .conf import RobotConfig
from .Foot_Trajectory import FootTrajectory
from .Footstep_Planner import Footstep


class WalkPlanner:
    def __init__(self, conf: RobotConfig = RobotConfig):
        self.conf = conf
        self.t = 0.0

    def plan(self, footsteps: List[Footstep]):
        conf = self.conf
        swing_trajectories = []
        for i in range(len(footsteps) - 2):
            start = [*footsteps[i].position, 0.0, footsteps[i].orientation[2]]
            target = [*footsteps[i + 2].position, 0.0, footsteps[i + 2].orientation[2]]
            swing_trajectories.append({
                "t0": self.t,
                "side": int(footsteps[i].side),  # the foot that swings during [t0, t0 + step_duration]
                "trajectory": FootTrajectory([self.t, self.t + conf.step_duration], start=start, target=target,
                                             step_height=conf.step_height, rise_ratio=conf.rise_ratio),
            })
            self.t += conf.step_duration
        return swing_trajectories
