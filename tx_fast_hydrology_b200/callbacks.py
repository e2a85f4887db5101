"""Callback protocol of the model object: six no-op hooks, same names and firing
points as the reference's `BaseCallback` (tx_fast_hydrology/callbacks.py:1-20):

    __on_step_start__ / __on_step_end__            around every step
    __on_simulation_start__ / __on_simulation_end__ around `simulate`
    __on_save_state__ / __on_load_state__          from `save_state` / `load_state`

Callbacks written against the reference (reading or mutating `model.o_t_next` /
`model.i_t_next` as numpy arrays) keep working: the model materialises its
device state on the host when they look at it and takes their changes back
before the next device launch.
"""


class BaseCallback:
    def __init__(self, *args, **kwargs):
        pass

    def __on_step_start__(self):
        return None

    def __on_step_end__(self):
        return None

    def __on_save_state__(self):
        return None

    def __on_load_state__(self):
        return None

    def __on_simulation_start__(self):
        return None

    def __on_simulation_end__(self):
        return None
