"""Timer and signature introspection (reference: misc_utilities.py:22-73)."""
import time
from inspect import signature

import numpy as np


class Timer:
    """wall-clock accumulators keyed by phase name (misc_utilities.py:22-59).  With
    `i_sync_context` the device is synchronised before a timer stops, as in the reference."""

    def __init__(self, i_sync_context=0):
        self.i_sync_context = i_sync_context
        self.timings = {}
        self.flags = {}

    def start(self, timer_key):
        if timer_key not in self.timings:
            self.timings[timer_key] = 0.
        self.flags[timer_key] = time.time()

    def stop(self, timer_key):
        if self.flags.get(timer_key) is None:
            raise ValueError('No time measurement in progress for timer ' + str(timer_key) + '.')
        if self.i_sync_context:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        self.timings[timer_key] += time.time() - self.flags[timer_key]
        self.flags[timer_key] = None

    def print_report(self):
        total = self.timings.get('total', sum(self.timings.values()))
        print('took ' + str(np.round(total / 60, 2)) + ' min.')
        print('Detailed computing times:')
        for key, value in self.timings.items():
            print(key + '\t' + str(np.round(100 * value / max(total, 1e-30), 0)) + '\t%\t' +
                  str(np.round(value, 1)) + ' \tsec')


def function_input_fields(function):
    """model-field names = argument names of a factory method (misc_utilities.py:63-73)"""
    input_fields = list(signature(function).parameters)
    for ign in ('self', 'GR', 'GRF'):
        if ign in input_fields:
            input_fields.remove(ign)
    return input_fields
