"""config.json reader in the reference's format (utils/config.py:30-50): lists become tuples, each top-level
section becomes an attribute bag."""
import json


class _Config(object):
    """Empty attribute bag; its __dict__ is what gets serialised to config.json."""
    pass


class LoadedRunConfig:
    def __init__(self, **entries):
        self.__dict__.update(entries)


def get_config_from_file(absolute_file_path):
    """Returns (model_config, train_config) from a config.json written by a reference (or this) run."""
    with open(absolute_file_path) as f:
        config_json = json.load(f)
    sections = {}
    for section, fields in config_json.items():
        sections[section] = LoadedRunConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in fields.items()})
    return sections['model'], sections['train']
