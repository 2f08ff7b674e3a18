"""Import shim (test infrastructure): skimage.util.shape.view_as_windows for
step=1 is numpy's sliding_window_view with the same output shape convention
(SURVEY.md section 8c)."""
from numpy.lib.stride_tricks import sliding_window_view


def view_as_windows(arr_in, window_shape, step=1):
    assert step == 1
    return sliding_window_view(arr_in, window_shape)
