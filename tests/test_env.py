"""Python restatement of the reference's test environment (/root/reference/test/test_env.jl:10-87), TEST INFRASTRUCTURE.

TestMDP: 3 observation patterns (bad / normal / good), an episode lasts max_time - 1 = 5 actions, the reward of a step is the
pattern's reward (-0.1, 0, +0.1) and is multiplied by -10 when the previous step ended in pattern 2 ("visiting the second state").
Optimal return 2.1 with the action sequence [2, 1, 2, 1, 3].  The environment protocol is the one
deepqlearning.jl_b200/solver.py documents (CommonRLInterface spelled in Python: reset / actions / observe / act / terminated)."""
import numpy as np


class TestMDP:
    __test__ = False          # not a pytest class

    def __init__(self, shape=(6,), stack=4, max_time=6, discount_factor=0.99, seed=7):
        rng = np.random.default_rng(seed)
        self.shape = tuple(shape)
        self.stack = 4                      # test_env.jl:33 passes the literal 4 here and `stack` as o_stack
        self.o_stack = int(stack)
        self.max_time = int(max_time)
        self.bad_state = rng.integers(1, 51, self.shape).astype(np.int32)
        self.normal_state = rng.integers(100, 151, self.shape).astype(np.int32)
        self.good_state = rng.integers(150, 201, self.shape).astype(np.int32)
        self._observation_space = [self.bad_state, self.normal_state, self.good_state]
        self._rewards = np.array([-0.1, 0.0, 0.1], np.float32)
        self.discount = float(np.float32(discount_factor))
        self.reset()

    # ---- POMDPs.jl pieces (test_env.jl:37-87) -------------------------------------------------------
    def actions(self):
        return [1, 2, 3, 4]

    def initialstate(self):
        return (np.ones(self.stack, np.int32), 1)

    def convert_s(self, s):
        """observation (W.., o_stack) of the reference == numpy (o_stack, ..reversed shape): the same memory image"""
        obs = np.zeros((self.o_stack,) + tuple(reversed(self.shape)), np.float32)
        for i in range(1, self.o_stack + 1):
            obs[i - 1] = self._observation_space[s[0][len(s[0]) - i] - 1].T      # Julia column-major array -> row-major transposed view
        return (obs / np.float32(255.0)).astype(np.float32)

    def gen(self, s, a):
        s_new = np.roll(s[0], -1)
        s_new[-1] = a if a < 4 else s_new[-2]
        return (s_new, s[1] + 1)

    def reward(self, s, a, sp):
        r = self._rewards[sp[0][-1] - 1]
        if s[0][-1] == 2:
            r = r * np.float32(-10)
        return float(r)

    def isterminal(self, s):
        return s[1] >= self.max_time

    # ---- CommonRLInterface protocol -----------------------------------------------------------------
    def reset(self):
        self.s = self.initialstate()

    def observe(self):
        return self.convert_s(self.s)

    def act(self, a):
        sp = self.gen(self.s, a)
        r = self.reward(self.s, a, sp)
        self.s = sp
        return r

    def terminated(self):
        return self.isterminal(self.s)


def evaluate(env, policy, action_fn, n_ep=100, max_steps=100):
    """test/runtests.jl:28-43"""
    avg_r = 0.0
    for _ in range(n_ep):
        r, step = 0.0, 0
        env.reset()
        while not env.terminated() and step < max_steps:
            r += env.act(action_fn(policy, env.observe()))
            step += 1
        avg_r += r
    return avg_r / n_ep
