"""results/{game}/{run_id}/... layout of the reference (paths.py:7-46), only what self-play writes or reads."""
import os

from custom_alphazero.config import ConfigGeneral, ConfigPath


def get_run_path(run_id: str) -> str:
    return os.path.join(ConfigPath.results_dir, ConfigGeneral.game, run_id)


def get_self_play_path(run_id: str) -> str:
    return os.path.join(get_run_path(run_id), ConfigPath.self_play_dir)


def get_self_play_iteration_path(run_id: str, iteration: int) -> str:
    return os.path.join(get_self_play_path(run_id), "iteration_{}".format(iteration))


def get_self_play_samples_path(run_id: str, iteration: int) -> str:
    return os.path.join(get_self_play_iteration_path(run_id, iteration), ConfigPath.samples_file)


def get_training_path(run_id: str) -> str:
    return os.path.join(get_run_path(run_id), ConfigPath.training_dir)


def get_evaluation_path(run_id: str) -> str:
    return os.path.join(get_run_path(run_id), ConfigPath.evaluation_dir)
