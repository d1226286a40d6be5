"""Test-side restatements of reference controllers that are NOT part of the product (SURVEY.md §2 row 12:
robot-specific controllers are out of scope) but are needed to replay the reference's own tests of the
in-scope dynamics path. Host logic only; they produce joint torques for step()."""
import math

import numpy as np


class QuadrupedTrottingController:
    """control/quadruped_control.rs:10-266 (QuadrupedTrottingController): trot gait scheduler, Raibert
    touchdown, 2-link inverse kinematics, joint PD. q/v are the flat vectors of build_quadruped."""

    def __init__(self, dt, target_x, default_foot_z):
        self.ticks = 0
        self.contact_phases = [(True, True, True, True), (False, True, True, False), (True, True, True, True),
                               (True, False, False, True)]
        self.dt = dt
        self.overlap_time = 0.1
        self.swing_time = 0.15
        self.z_clearance = 0.25
        self.default_foot_z = default_foot_z
        self.default_stance = [np.array([0.0, 0.0, default_foot_z]) for _ in range(4)]
        self.foot_locations = [s.copy() for s in self.default_stance]
        self.vx = 0.0
        self.target_x = target_x
        self.l_leg = 1.0

    # ---- tick bookkeeping (:218-266) --------------------------------------------------------
    def overlap_ticks(self):
        return int(self.overlap_time / self.dt)

    def swing_ticks(self):
        return int(self.swing_time / self.dt)

    def stance_ticks(self):
        return 2 * self.overlap_ticks() + self.swing_ticks()

    def period_ticks(self):
        return 2 * self.overlap_ticks() + 2 * self.swing_ticks()

    def phase_ticks_vec(self):
        o, s = self.overlap_ticks(), self.swing_ticks()
        return [o, s, o, s]

    def phase_index(self, ticks):
        phase_time = ticks % self.period_ticks()
        total = 0
        for i, t in enumerate(self.phase_ticks_vec()):
            total += t
            if phase_time < total:
                return i
        raise AssertionError("should not reach this")

    def subphase_ticks(self, ticks):
        phase_time = ticks % self.period_ticks()
        total = 0
        for t in self.phase_ticks_vec():
            total += t
            if phase_time < total:
                return phase_time + t - total
        raise AssertionError("should not reach this")

    # ---- foot targets (:107-160) --------------------------------------------------------------
    def next_stance_foot_location(self, leg):
        loc = self.foot_locations[leg]
        v = np.array([-self.vx, 0.0, 1.0 / 0.02 * (self.default_foot_z - loc[2])])
        return loc + v * self.dt

    def next_swing_foot_location(self, swing_ticks, leg):
        swing_proportion = swing_ticks / self.swing_ticks()
        assert 0.0 <= swing_proportion <= 1.0
        loc = self.foot_locations[leg]
        height_proportion = (swing_ticks + 1) / self.swing_ticks()
        if height_proportion < 0.5:
            swing_height = self.z_clearance * height_proportion / 0.5
        else:
            swing_height = self.z_clearance * (1.0 - (height_proportion - 0.5) / 0.5)
        z_vector = np.array([0.0, 0.0, swing_height + self.default_foot_z])
        delta_px = 0.5 * self.stance_ticks() * self.dt * self.vx
        touchdown = self.default_stance[leg] + np.array([delta_px, 0.0, 0.0])
        time_left = self.dt * self.swing_ticks() * (1.0 - swing_proportion)
        v = ((touchdown - loc) / time_left) * np.array([1.0, 1.0, 0.0])
        return loc * np.array([1.0, 1.0, 0.0]) + z_vector + v * self.dt

    def step_gait(self):
        modes = self.contact_phases[self.phase_index(self.ticks)]
        return [self.next_stance_foot_location(leg) if modes[leg]
                else self.next_swing_foot_location(self.subphase_ticks(self.ticks), leg) for leg in range(4)]

    @staticmethod
    def inverse_kinematics(foot_locations, l_leg):  # :163-189
        out = []
        for f in foot_locations:
            x, z = f[0], f[2]
            l1 = l2 = l_leg / 2.0
            cos_theta2 = (x * x + z * z - l1 * l1 - l2 * l2) / (2.0 * l1 * l2)
            theta2 = math.acos(max(-1.0, min(1.0, cos_theta2))) if abs(cos_theta2) <= 1.0 + 1e-15 else float("nan")
            theta1 = math.atan2(x, -z) - math.atan2(l2 * math.sin(theta2), l1 + l2 * math.cos(theta2))
            if theta1 > 0.0:
                theta1 -= math.pi
            out.append((theta1, theta2))
        return out

    def control(self, q, v):  # :28-71
        x = q[4]
        dx = x - self.target_x
        self.vx = -math.copysign(1.0, dx) * min(abs(dx) * 10.0, 1.0)
        tau = np.zeros(14)
        self.foot_locations = self.step_gait()
        angles = self.inverse_kinematics(self.foot_locations, self.l_leg)
        self.ticks += 1
        kp, kd = 150.0, 10.0
        for leg, (hip_angle, knee_angle) in enumerate(angles):
            hip = leg * 2  # scalar joint index: q[7 + hip], v[6 + hip]
            tau[6 + hip] = kp * (hip_angle - q[7 + hip]) + kd * -v[6 + hip]
            tau[6 + hip + 1] = kp * (knee_angle - q[7 + hip + 1]) + kd * -v[6 + hip + 1]
        return tau


def quadruped_initial_state(default_z=0.8, initial_x=-1.2, l_leg=1.0):
    """quadruped_trot_to_position, control/quadruped_control.rs:417-452"""
    q = np.zeros(15)
    q[3] = 1.0
    q[4], q[6] = initial_x, default_z
    angles = QuadrupedTrottingController.inverse_kinematics([np.array([0.0, 0.0, -default_z])] * 4, l_leg)
    for leg, (hip, knee) in enumerate(angles):
        q[7 + 2 * leg] = hip
        q[7 + 2 * leg + 1] = knee
    return q, np.zeros(14)
