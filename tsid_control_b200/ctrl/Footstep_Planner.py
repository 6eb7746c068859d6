"""Footstep, Support and FootstepPlanner — counterpart of ref:ctrl/Footstep_Planner.py:4-125.  Same classes,
methods and arithmetic; the reference's import-time demo and matplotlib plot (:127-186) are not reproduced.
"""
from typing import List

import numpy as np


class Footstep:
    def __init__(self, position: np.ndarray, orientation: np.ndarray, side: int):
        self.frame = np.eye(3)
        self.position = np.array(position)
        self.orientation = np.array(orientation)
        self.side = side  # 0 left, 1 right
        self.update_frame()

    def update_frame(self):
        self.frame[:2, 2] = self.position
        self.frame[:2, :2] = self.rotation_matrix(self.orientation)

    def rotation_matrix(self, orientation) -> np.ndarray:
        sz, cz = np.sin(orientation[2]), np.cos(orientation[2])
        return np.array([[cz, -sz], [sz, cz]])

    def transform(self, point) -> np.ndarray:
        return self.frame[:2, :2] @ point + self.frame[:2, 2]

    def __repr__(self):
        return f"Footstep(position={self.position}, orientation={self.orientation})"


class Support:
    def __init__(self, contacts: List[Footstep], foot_width: float, foot_length: float, start_time: float = 0.0):
        self.contacts = contacts
        self.is_double_support = len(contacts) == 2
        self.foot_width = foot_width
        self.foot_length = foot_length
        self.start_time = start_time

    def get_support_polygon(self) -> List[np.ndarray]:
        """Vertices in the order of ref:ctrl/Footstep_Planner.py:48-66."""
        L, W = self.foot_length / 2, self.foot_width / 2
        polygon = []
        for contact in self.contacts:
            if contact.side == 0:
                corners = ([-L, W], [-L, -W], [L, -W], [L, W])
            else:
                corners = ([L, -W], [L, W], [-L, W], [-L, -W])
            polygon.extend(contact.transform(c) for c in corners)
        return polygon


class FootstepPlanner:
    def __init__(self, step_width, step_length):
        self.step_width = step_width
        self.step_length = step_length

    def add_step(self, dx, dy, side: int, pos: np.ndarray) -> Footstep:
        tangent = np.array([dx, dy])
        tangent /= np.linalg.norm(tangent)
        normal = np.array([-tangent[1], tangent[0]])
        position = pos + tangent * (self.step_length / 2) + normal * (self.step_width / 2 * (1 if side == 0 else -1))
        orientation = np.array([0, 0, np.arctan2(dy, dx)])
        return Footstep(position=position, orientation=orientation, side=side)

    def plan(self, path: List[np.ndarray], init_supports: List[Footstep]) -> List[Footstep]:
        """ref:ctrl/Footstep_Planner.py:92-125: a new step every time the accumulated path length reaches
        step_length, feet alternating; the final step(s) close the path in double support."""
        footsteps = []
        footsteps.extend(init_supports)
        side = init_supports[-1].side
        distance = 0.0
        for i in range(len(path) - 1):
            dx, dy = path[i + 1] - path[i]
            distance += np.linalg.norm([dx, dy])
            if distance >= self.step_length:
                side = not side
                footsteps.append(self.add_step(dx, dy, side, path[i]))
                distance = 0.0
        side = not side
        footsteps.append(self.add_step(dx, dy, side, path[-1]))
        if distance > 0:
            side = not side
            footsteps.append(self.add_step(dx, dy, side, path[-1]))
        return footsteps
