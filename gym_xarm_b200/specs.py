"""Task table of the five hot-path envs: ids, dims, default configs, TimeLimit."""
from dataclasses import dataclass, field

TASK_IDS = {"reach": 0, "pick_and_place": 1, "stack_tower": 2, "push_with_door": 3, "handover": 4}
REWARD_IDS = {"sparse": 0, "dense": 1, "dense_o2g": 2, "dense_diff": 3}
GOAL_IDS = {"air": 0, "ground": 1}


@dataclass(frozen=True)
class TaskSpec:
    name: str
    task: int
    ref_class: str           # class name in the reference
    ref_file: str            # file in /root/reference/gym_xarm/envs
    act_dim: int
    max_episode_steps: int   # TimeLimit / env._max_episode_steps
    distance_threshold: float
    reward_types: tuple
    default_config: dict = field(default_factory=dict)

    def dims(self, num_obj=1):
        if self.name == "reach":
            return self.act_dim, 8, 3
        if self.name == "pick_and_place":
            return self.act_dim, 8 + 16 * num_obj, 3 * num_obj
        if self.name == "stack_tower":
            return self.act_dim, 55, 9
        if self.name == "push_with_door":
            return self.act_dim, 25, 3
        return self.act_dim, 13 * num_obj + 16, 3 * num_obj


SPECS = {
    "reach": TaskSpec("reach", 0, "XarmReachEnv", "xarm_reach.py", 4, 25, 0.05, ("sparse", "dense", "dense_diff"),
                      {"GUI": False, "reward_type": "sparse"}),
    "pick_and_place": TaskSpec("pick_and_place", 1, "XarmPickAndPlace", "xarm_pick_and_place.py", 4, 50, 0.05,
                               ("sparse", "dense", "dense_o2g"),
                               {"GUI": False, "reward_type": "sparse", "num_obj": 1, "goal_shape": "air",
                                "init_grasp_rate": 0.0, "goal_ground_rate": 0.0}),
    "stack_tower": TaskSpec("stack_tower", 2, "XarmStackTowerEnv", "xarm_stack_tower.py", 8, 50, 0.09, ("sparse", "dense"),
                            {"GUI": False, "reward_type": "sparse"}),
    "push_with_door": TaskSpec("push_with_door", 3, "XarmPushWithDoorEnv", "xarm_push_with_door.py", 6, 50, 0.03,
                               ("sparse", "dense"), {"GUI": False, "reward_type": "sparse"}),
    "handover": TaskSpec("handover", 4, "XarmHandover", "xarm_handover.py", 8, 100, 0.05, ("sparse", "dense"),
                         {"GUI": False, "reward_type": "sparse", "num_obj": 1, "goal_shape": "ground",
                          "same_side_rate": 0.5, "use_stand": False}),
}


def normalize_config(spec, config):
    """Merge the reference-style `config` dict over the task defaults and validate it like the reference would."""
    cfg = dict(spec.default_config)
    if config:
        cfg.update(config)
    rt = cfg.get("reward_type", "sparse")
    if rt not in REWARD_IDS or rt not in spec.reward_types:
        # [REF xarm_pick_and_place.py:189-190] unknown reward types raise NotImplementedError (D6 for the broken ones)
        raise NotImplementedError(f"reward_type {rt!r} is not implemented for {spec.ref_class}")
    gs = cfg.get("goal_shape", "air")
    if gs == "any":  # test.py passes goal_shape='any' to Handover: anything but 'ground' keeps the sampled height
        gs = "air"
    if gs not in GOAL_IDS:
        raise ValueError(f"goal_shape {gs!r}")
    cfg["goal_shape"] = gs
    n = int(cfg.get("num_obj", 1))
    if spec.name == "stack_tower":
        n = 3
    elif spec.name == "push_with_door":
        n = 1
    elif spec.name == "reach":
        n = 0
    cfg["num_obj"] = n
    return cfg
