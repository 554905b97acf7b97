"""Development convenience, NOT part of the product (the reference's command line, main.py:292-304, is out of
scope -- SURVEY section 2 row 11): run macjd_b200.main.run from a shell.

  python tools/train_cli.py [--config NAME --config-dir DIR] [--sim-config YAML] [--n-envs N]
                            [--total-env-steps K] [--pipeline] [--no-tensorboard]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from macjd_b200 import main as M
    ap = argparse.ArgumentParser(description="QMix / MP-DQN training on the batched device-resident path")
    ap.add_argument("--config", default=None, help="name of a YAML file under --config-dir (reference format); default: built-in defaults")
    ap.add_argument("--config-dir", default="config")
    ap.add_argument("--sim-config", default=None, help="scenario YAML (reference format); default: the reference's default scenario")
    ap.add_argument("--n-envs", type=int, default=1024)
    ap.add_argument("--total-env-steps", type=int, default=None)
    ap.add_argument("--train-steps-per-rollout", type=int, default=None)
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--no-tensorboard", action="store_true")
    a = ap.parse_args()
    cfg = M.load_config(a.config, a.config_dir) if a.config else M.default_config()
    if a.total_env_steps is not None:
        cfg.total_env_steps = a.total_env_steps
    if a.train_steps_per_rollout is not None:
        cfg.train_steps_per_rollout = a.train_steps_per_rollout
    kw = dict(writer=False if a.no_tensorboard else None, pipeline=a.pipeline)
    if a.sim_config is None:
        from macjd_b200.simulation.scenario import default_spec
        M.run(cfg, spec=default_spec(a.n_envs), **kw)
    else:
        M.run(cfg, n_envs=a.n_envs, sim_config_path=a.sim_config, **kw)
