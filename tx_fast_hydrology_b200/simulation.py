"""`CheckPoint` callback (tx_fast_hydrology/simulation.py:169-211): saves the model state once
when `model.datetime >= checkpoint_time` and fans save/load out to sibling callbacks.  The
sub-basin orchestration of simulation.py (`AsyncSimulation`) is host glue outside the routing hot
path and is not rebuilt here."""
import datetime
import logging

from .callbacks import BaseCallback

logger = logging.getLogger(__name__)


class CheckPoint(BaseCallback):
    def __init__(self, model, checkpoint_time=None, timedelta=None):
        self.model = model
        if checkpoint_time is None:
            if timedelta is None:
                raise ValueError('Either `checkpoint_time` or `timedelta` must not be `None`.')
            checkpoint_time = model.datetime + datetime.timedelta(seconds=timedelta)
        self.checkpoint_time = checkpoint_time
        self.timedelta = timedelta
        self.model_saved = False

    def __on_simulation_start__(self):
        if self.timedelta is None:
            return None
        self.set_checkpoint(self.model.datetime + datetime.timedelta(seconds=self.timedelta))

    def __on_step_end__(self):
        if (self.model.datetime >= self.checkpoint_time) and (not self.model_saved):
            self.model.save_state()
            self.model_saved = True

    def __on_save_state__(self):
        for _, callback in self.model.callbacks.items():
            if hasattr(callback, 'save_state'):
                callback.save_state()

    def __on_load_state__(self):
        for _, callback in self.model.callbacks.items():
            if hasattr(callback, 'load_state'):
                callback.load_state()

    def set_checkpoint(self, checkpoint_time):
        logger.info(f'Setting checkpoint time to {checkpoint_time}')
        self.checkpoint_time = checkpoint_time
        self.model_saved = False
