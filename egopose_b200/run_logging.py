"""Run logging the training scripts call (observability only):

  create_logger(path, file_handle)   utils/logger.py:5-26 contract: a logging.Logger that prints the bare message on the
                                     console and appends '[time] message' to ``path``
  Logger(log_dir).scalar_summary     utils/tb_logger.py:26-42 contract.  The reference writes TF1 event files; TensorFlow is
                                     not a dependency of the fused path, so scalars go to ``<log_dir>/scalars.jsonl`` (one
                                     {"tag", "value", "step", "time"} object per line) - convertible to any dashboard format.
"""
import json
import logging
import os
import time


def create_logger(filename, file_handle=True):
    log = logging.getLogger(filename)
    log.propagate = False
    log.setLevel(logging.DEBUG)
    if not any(isinstance(h, logging.StreamHandler) and not isinstance(h, logging.FileHandler) for h in log.handlers):
        console = logging.StreamHandler()
        console.setLevel(logging.INFO)
        console.setFormatter(logging.Formatter('%(message)s'))
        log.addHandler(console)
    if file_handle and not any(isinstance(h, logging.FileHandler) for h in log.handlers):
        os.makedirs(os.path.dirname(filename) or '.', exist_ok=True)
        fh = logging.FileHandler(filename, mode='a')
        fh.setLevel(logging.DEBUG)
        fh.setFormatter(logging.Formatter('[%(asctime)s] %(message)s'))
        log.addHandler(fh)
    return log


class Logger:
    """scalar logger with the reference's tb_logger interface (ego_mimic.py:128-131)"""

    def __init__(self, log_dir, name=None):
        self.name = name
        self.dir = os.path.join(log_dir, name) if name else log_dir
        os.makedirs(self.dir, exist_ok=True)
        self.path = os.path.join(self.dir, 'scalars.jsonl')
        self._f = open(self.path, 'a')

    def scalar_summary(self, tag, value, step):
        self._f.write(json.dumps({'tag': tag, 'value': float(value), 'step': int(step), 'time': time.time()}) + '\n')
        self._f.flush()

    def image_summary(self, tag, images, step):
        raise NotImplementedError('image summaries are out of scope of the fused path')

    def histo_summary(self, tag, values, step, bins=1000):
        raise NotImplementedError('histogram summaries are out of scope of the fused path')
