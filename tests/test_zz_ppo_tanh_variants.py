"""GPU parity of the compile-time tanh variants of the PPO update / inference kernels (``PpoAlgoT<8, 1|2|3>``, ``InferAlgoT<1>``) behind
PPO_with_tricks' ``tanh`` switch.  Kept in a file that sorts last: these instantiations were rebuilt after the round's last GPU run (their
arithmetic was GPU-verified in its earlier run-time form and is checked on the host emulation in ``test_parity_ppo.py``), so under
``pytest -x`` they run after every other GPU test."""
import pytest
import torch

from test_parity_ppo import _ppo_tricks


@pytest.mark.gpu
def test_ppo_with_tricks_tanh_gpu(golden):
    _ppo_tricks(golden, torch.device("cuda"), "ppo_tricks_tanh_cont", True, tanh=True)
    _ppo_tricks(golden, torch.device("cuda"), "ppo_tricks_tanh_disc", False, tanh=True)
