"""Where a run keeps its files: the reference's results/{game}/{run_id}/... layout (its paths.py), limited to
what self-play writes or reads.  All helpers are thin joins under one root so the layout lives in one table."""
import os

from custom_alphazero.config import ConfigGeneral, ConfigPath

_SUBDIRS = {"self_play": "self_play_dir", "training": "training_dir", "evaluation": "evaluation_dir"}


def _under(run_id, kind=None, *more):
    parts = [ConfigPath.results_dir, ConfigGeneral.game, run_id]
    if kind is not None:
        parts.append(getattr(ConfigPath, _SUBDIRS[kind]))
    return os.path.join(*parts, *more)


def _numbered(iteration, prefix="iteration", sep="_"):
    return f"{prefix}{sep}{iteration}"


def get_run_path(run_id):
    return _under(run_id)


def get_self_play_path(run_id):
    return _under(run_id, "self_play")


def get_training_path(run_id):
    return _under(run_id, "training")


def get_evaluation_path(run_id):
    return _under(run_id, "evaluation")


def get_self_play_iteration_path(run_id, iteration):
    return _under(run_id, "self_play", _numbered(iteration))


def get_self_play_samples_path(run_id, iteration):
    return _under(run_id, "self_play", _numbered(iteration), ConfigPath.samples_file)


def get_evaluation_iteration_path(run_id, iteration):
    return _under(run_id, "evaluation", _numbered(iteration))
