"""Scene shells: same names and attributes as the reference's scene_bases.py / scene_stadium.py.

The reference's Scene/World push gravity, ERP, solver iterations and sub-steps into pybullet
(scene_bases.py:60-76) and step it; here those numbers live in spec.SceneSpec and are baked into the
model tables, and ``global_step`` is a no-op because the env step is one fused kernel launch.
"""
from ...spec import SceneSpec


class World:
    def __init__(self, gravity, timestep, frame_skip, num_solver_iterations=5):
        self.gravity, self.timestep, self.frame_skip = gravity, timestep, frame_skip
        self.numSolverIterations = num_solver_iterations

    def clean_everything(self):
        pass

    def step(self, frame_skip):
        pass


class Scene:
    "A base class for single- and multiplayer scenes"
    multiplayer = False

    def __init__(self, scene_spec: SceneSpec):
        self.spec = scene_spec
        self.timestep = scene_spec.timestep
        self.frame_skip = scene_spec.frame_skip
        self.dt = self.timestep * self.frame_skip
        self.cpp_world = World(scene_spec.gravity, self.timestep, self.frame_skip, scene_spec.num_solver_iterations)
        self.test_window_still_open = True
        self.human_render_detected = False
        self.multiplayer_robots = {}

    def test_window(self):
        self.human_render_detected = True
        return self.test_window_still_open

    def actor_introduce(self, robot):
        pass

    def actor_is_active(self, robot):
        return not self.multiplayer

    def episode_restart(self, bullet_client=None):
        self.cpp_world.clean_everything()

    def global_step(self):
        self.cpp_world.step(self.frame_skip)


class SingleRobotEmptyScene(Scene):
    multiplayer = False


class StadiumScene(Scene):
    multiplayer = False
    zero_at_running_strip_start_line = True
    stadium_halflen = 105 * 0.25
    stadium_halfwidth = 50 * 0.25
    stadiumLoaded = 0

    def episode_restart(self, bullet_client=None):
        Scene.episode_restart(self, bullet_client)
        self.stadiumLoaded = 1
