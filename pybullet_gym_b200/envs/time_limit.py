"""gym.wrappers.TimeLimit semantics of the gym era the reference targets (0.9.6 - 0.10.x)."""


class TimeLimit:
    def __init__(self, env, max_episode_steps=None):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.metadata = getattr(env, "metadata", {})

    @property
    def unwrapped(self):
        return self.env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._max_episode_steps is not None and self._elapsed_steps >= self._max_episode_steps:
            done = True
        return obs, reward, done, info

    def reset(self):
        self._elapsed_steps = 0
        return self.env.reset()

    def seed(self, seed=None):
        return self.env.seed(seed)

    def render(self, mode="human", **kw):
        return self.env.render(mode, **kw)

    def close(self):
        return self.env.close()
