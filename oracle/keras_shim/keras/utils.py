"""keras.utils subset: there is no network, so `get_file` answers with a path that does not exist (the reference then
skips loading, ckpt_loader is only reached through explicit local paths) and the progress bar is silent."""
import os


def get_file(fname=None, origin=None, **kwargs):
    return os.path.join("/nonexistent-keras-cache", str(fname or os.path.basename(str(origin))))


class Progbar:
    def __init__(self, target, **kwargs):
        self.target = target

    def update(self, current, **kwargs):
        pass
