"""Config reading -- same semantics as nlt/util/io.py:40-44."""
from configparser import ConfigParser
from os.path import exists, join, dirname, abspath


def read_config(path):
    if not exists(path):   # allow bare names of the shipped configs
        alt = join(dirname(dirname(abspath(__file__))), 'config', path)
        if exists(alt):
            path = alt
    config = ConfigParser()
    with open(path, 'r') as h:
        config.read_file(h)
    return config


def make_config(**overrides):
    """A [DEFAULT]-only ConfigParser with the shipped dragon_specular keys
    (nlt/config/dragon_specular.ini), overridden by keyword."""
    base = dict(
        dataset='nlt', no_batch='False', bs='4', model='nlt', loss='l2', lr='1e-3', mgm='-1', epochs='100',
        imh='512', imw='512', uvh='512', uvw='512', linear_space='False', use_obs='True',
        skip_connect_base='True', depth0='16', depth='256', kernel='2', stride='2', norm='None',
        act='leakyrelu', pool='None')
    base.update({k: str(v) for k, v in overrides.items()})
    config = ConfigParser()
    config.read_dict({'DEFAULT': base})
    return config
