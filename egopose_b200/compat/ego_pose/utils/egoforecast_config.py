"""ego_pose/utils/egoforecast_config.py mirror: the same attribute bag, reading config/egoforecast/<id>.yml"""
from egopose_b200.config import Config as _Config


class Config(_Config):
    def __init__(self, cfg_id=None, create_dirs=False, cfg_dict=None, base_dir='results'):
        super().__init__(cfg_id, create_dirs=create_dirs, cfg_dict=cfg_dict, task='egoforecast', base_dir=base_dir)
