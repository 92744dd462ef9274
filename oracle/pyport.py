"""Pure-Python single-env port of the reference's step/reset with its dict API — TEST
INFRASTRUCTURE (see oracle/__init__.py).  Independent of cc_oracle.c: plain lists and dicts,
float64 rewards, numpy only for the PCG64 generator and the float32 observation arrays.

It exists (a) to check the single-env facade on the GPU box, where the reference itself is not
available, with arbitrary action dicts, and (b) as the Python-speed CPU baseline of
``bench.py --impl reference`` (the reference is a pure-Python loop).  Pinned against the
unmodified reference in tests/test_pyport.py.  Citations: /root/reference/src/collectivecrossing/.
"""

from __future__ import annotations

import numpy as np

from collectivecrossing_b200.utils.geometry import calculate_tram_boundaries

MOVES = ((1, 0), (0, 1), (-1, 0), (0, -1), (0, 0))  # actions.py:18-24


class PyEnv:
    def __init__(self, config):
        self.c = config
        tb = calculate_tram_boundaries(config)
        self.TL, self.TR, self.DL, self.DR = tb.tram_left, tb.tram_right, tb.tram_door_left, tb.tram_door_right
        self.DC = (self.DL + self.DR) // 2
        self.ids = [f"boarding_{i}" for i in range(config.num_boarding_agents)] + [
            f"exiting_{i}" for i in range(config.num_exiting_agents)]
        self.boarding = {i: k < config.num_boarding_agents for k, i in enumerate(self.ids)}
        self.reward_name = config.reward_config.get_reward_function_name()
        self.term_name = config.terminated_config.get_terminated_function_name()
        self.max_steps = config.truncated_config.max_steps
        self.rng = None
        self.pos, self.active, self.term, self.trunc, self.steps = {}, {}, {}, {}, 0

    # ---- geometry (collectivecrossing.py:509-563) ----
    def valid(self, x, y):
        c = self.c
        if not (0 <= x <= c.width and 0 <= y <= c.height):
            return False
        if y == c.division_y and not (self.DL < x < self.DR):
            return False
        if y >= c.division_y and not (self.TL < x < self.TR):
            return False
        return True

    def occupied(self, x, y, skip=None):
        return any(i != skip and self.active[i] and self.pos[i] == (x, y) for i in self.pos)

    def arrived(self, i):
        y = self.pos[i][1]
        return y == (self.c.boarding_destination_area_y if self.boarding[i] else self.c.exiting_destination_area_y)

    def in_tram(self, i):
        x, y = self.pos[i]
        return y >= self.c.division_y and self.TL <= x <= self.TR

    def at_door(self, i):
        x, y = self.pos[i]
        return y == self.c.division_y and (x == self.DL - 1 or x == self.DR + 1)

    # ---- reset (collectivecrossing.py:91-159) ----
    def reset(self, seed=None):
        if seed is not None or self.rng is None:
            self.rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        c, g = self.c, self.rng
        self.pos, self.active, self.term, self.trunc, self.steps = {}, {}, {}, {}, 0
        for i in self.ids:
            while True:
                if self.boarding[i]:
                    x, y = int(g.integers(0, c.width)), int(g.integers(0, c.division_y))
                    ok = self.valid(x, y) and not self.occupied(x, y) and not (self.DL <= x <= self.DR and y == c.division_y - 1)
                else:
                    x, y = int(g.integers(self.TL, self.TR + 1)), int(g.integers(c.division_y, c.height))
                    ok = self.valid(x, y) and not self.occupied(x, y)
                if ok:
                    self.pos[i], self.active[i], self.term[i], self.trunc[i] = (x, y), True, False, False
                    break
        return {i: self.observation(i) for i in self.ids}, {i: {"agent_type": self.kind(i)} for i in self.ids}

    def kind(self, i):
        return "boarding" if self.boarding[i] else "exiting"

    # ---- observation (observations.py:43-94) ----
    def observation(self, i):
        row = [*self.pos[i], self.DC, self.c.division_y, self.DL, self.DR]
        for j in self.ids:
            row += [-1, -1, -1, -1] if j == i else [*self.pos[j], 0 if self.boarding[j] else 1, 1 if self.active[j] else 0]
        return np.array(row, dtype=np.float32)

    # ---- rewards (rewards.py:41-182) ----
    def reward(self, i):
        p, (x, y), D = self.c.reward_config, self.pos[i], self.c.division_y
        if self.reward_name == "default":
            if self.boarding[i]:
                if self.arrived(i):
                    return p.boarding_destination_reward
                if self.at_door(i):
                    return p.tram_door_reward
                if self.in_tram(i):
                    return p.tram_area_reward
                return -(abs(x - self.DC) + (D - y)) * p.distance_penalty_factor
            if self.arrived(i):
                return p.boarding_destination_reward
            if not self.in_tram(i):
                return p.tram_area_reward
            return (abs(x - self.DC) + (y - D)) * p.distance_penalty_factor
        if self.reward_name == "simple_distance":
            goal = self.c.boarding_destination_area_y if self.boarding[i] else self.c.exiting_destination_area_y
            return -abs(y - goal) * p.distance_penalty_factor
        if self.reward_name == "binary":
            return p.no_goal_reward  # the goal comparison of the reference can never be true
        return p.step_penalty

    # ---- step (collectivecrossing.py:161-261) ----
    def step(self, actions):
        self.steps += 1
        alive = {i: not (self.term[i] or self.trunc[i]) for i in self.ids}
        for i, a in actions.items():
            if i not in self.pos:
                raise ValueError(f"Unknown agent ID: {i} in action_dict.")
            if a not in (0, 1, 2, 3, 4):
                raise ValueError(f"Invalid action: {a} for agent {i}. Valid actions are: [0, 1, 2, 3, 4]")
            if self.active[i]:
                x, y = self.pos[i][0] + MOVES[a][0], self.pos[i][1] + MOVES[a][1]
                if self.valid(x, y) and not self.occupied(x, y, skip=i):
                    self.pos[i] = (x, y)
        for i in self.ids:
            if self.active[i] and self.arrived(i):
                self.active[i] = False
        rewards = {i: float(self.reward(i)) for i in self.ids if alive[i]}
        everyone = all(self.arrived(i) for i in self.ids)
        terminateds = {i: (everyone if self.term_name == "all_at_destination" else self.arrived(i)) for i in self.ids}
        truncateds = {i: self.steps >= self.max_steps for i in self.ids if alive[i]}
        fresh = set()
        for i in self.ids:
            if terminateds[i] and not self.term[i]:
                self.term[i] = True
                fresh.add(i)
        for i, v in truncateds.items():
            if v and not self.trunc[i]:
                self.trunc[i] = True
                fresh.add(i)
        shown = [i for i in self.ids if not (self.term[i] or self.trunc[i]) or i in fresh]
        obs = {i: self.observation(i) for i in shown}
        infos = {i: {"agent_type": self.kind(i), "in_tram_area": self.in_tram(i), "at_door": self.at_door(i),
                     "active": self.active[i], "at_destination": self.arrived(i)} for i in shown}
        terminateds["__all__"] = all(terminateds.values()) if terminateds else False
        truncateds["__all__"] = all(truncateds.values()) if truncateds else False
        return obs, rewards, terminateds, truncateds, infos

    @property
    def possible_agents(self):
        return list(self.ids)

    @property
    def agents(self):
        return [i for i in self.ids if not (self.term[i] or self.trunc[i])]

    # ---- greedy / waiting at epsilon 0 (baseline_policies/*.py) ----
    def move_ok(self, i, a):
        if a == 4:
            return True
        x, y = self.pos[i][0] + MOVES[a][0], self.pos[i][1] + MOVES[a][1]
        return self.valid(x, y) and not self.occupied(x, y, skip=i)

    def greedy_action(self, i):
        (x, y), D, dc = self.pos[i], self.c.division_y, self.DC
        sgn = lambda v: (v > 0) - (v < 0)  # noqa: E731
        dx = dy = 0
        if self.boarding[i]:
            if y < D:
                if y == D - 1 and x != dc:
                    dx = sgn(dc - x)
                else:
                    dy = 1
                pref = [0, 1, 2, 3] if x < dc else [2, 1, 0, 3] if x > dc else [1, 0, 2, 3]
            else:
                dy, pref = sgn(self.c.boarding_destination_area_y - y), [1, 0, 2, 3]
        else:
            if y > D:
                if y == D + 1 and x != dc:
                    dx = sgn(dc - x)
                else:
                    dy = -1
                pref = [0, 3, 2, 1] if x < dc else [2, 3, 0, 1] if x > dc else [3, 0, 2, 1]
            else:
                dy, pref = sgn(self.c.exiting_destination_area_y - y), [3, 0, 2, 1]
        want = {(1, 0): 0, (0, 1): 1, (-1, 0): 2, (0, -1): 3, (0, 0): 4}[(dx, dy)]
        if self.move_ok(i, want):
            return want
        return next((a for a in pref if self.move_ok(i, a)), 4)

    def waiting_action(self, i):
        if self.boarding[i] and not self.in_tram(i) and any(
                not self.boarding[j] and not (self.term[j] or self.trunc[j]) and not self.arrived(j) for j in self.ids):
            return 4
        return self.greedy_action(i)

    def policy_actions(self, policy):
        fn = {"greedy": self.greedy_action, "waiting": self.waiting_action}[policy]
        return {i: fn(i) for i in self.agents if self.active[i]}
