"""Bit-pattern parity (signed zeros included) of the CUDA path against the oracle on the stack scene."""
import numpy as np
import pytest

import refdrv
import scenes

pytestmark = pytest.mark.gpu


def test_stack_state_bit_patterns_equal_the_oracle(pkg):
    desc = scenes.stack()
    batch = pkg.Batch(pkg.Scene(desc), n_worlds=4, device=0)
    batch.set_scene_forces(desc)
    oracle = refdrv.RefWorld("strict" if refdrv.available("strict") else "port").load(desc)
    for frame in range(30):
        batch.step()
        oracle.step()
        got = np.ascontiguousarray(batch.state()[0, :, :13])
        want = np.ascontiguousarray(oracle.state()[:, :13])
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), "frame %d: bit patterns differ (signed zero?)" % frame
    batch.close()
