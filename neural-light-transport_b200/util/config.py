"""nlt/util/config.py:15-22."""


def config2dict(config):
    return {k: config.get('DEFAULT', k) for k in config['DEFAULT']}
