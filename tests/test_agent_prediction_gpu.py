"""Device-side update_prediction (peanut_b200/agent_prediction.py, pn_target_pred) against the oracle pinned to the
reference's Agent_State.update_prediction: bit-exact for the glue (stamp, window embedding, bounds, unexplored mask) on an
injected prediction, and within the stage-C tolerance end to end."""
import numpy as np
import pytest
import torch

from oracle import agent_prediction as O
from oracle import prednet as OC
from peanut_b200 import agent_prediction, prediction

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("full,local,win,seed,goal", [(96, 48, 72, 1, 2), (96, 48, 96, 2, 0), (120, 60, 72, 3, 5), (96, 48, 40, 4, 3)])
def test_glue_bit_exact_given_the_prediction(full, local, win, seed, goal):
    fm, lm, lmb = O.synth_state(seed, full, local)
    fm_o = fm.copy()
    ref = O.update_prediction(fm_o, lm, lmb, goal, O.fake_prediction, win)
    # the same prediction, computed by the oracle's stand-in model on the stamped window, injected on the device
    x1 = 0 if win == full else full // 2 - win // 2
    preds = O.fake_prediction(fm_o[:, x1:x1 + win, x1:x1 + win])
    seg = prediction.init_segmentor(prediction._default_cfg(14, 6), device="cuda:0", precision="tf32",
                                    state_dict=OC.synth_state_dict(14, 6, seed=0))
    fm_d = torch.from_numpy(fm).cuda()
    got = agent_prediction.update_prediction(fm_d, torch.from_numpy(lm).cuda(), lmb, goal, seg, win,
                                             object_preds=torch.from_numpy(preds).cuda())
    assert got.dtype == ref.dtype and np.array_equal(got, ref)
    assert np.array_equal(fm_d.cpu().numpy(), fm_o)


def test_end_to_end_with_the_network():
    full, local, win, goal = 96, 48, 64, 1
    fm, lm, lmb = O.synth_state(7, full, local)
    sd = OC.synth_state_dict(14, 6, seed=0)
    model = OC.build(sd)
    ref = O.update_prediction(fm.copy(), lm, lmb, goal, lambda x: OC.get_prediction(model, np.ascontiguousarray(x)), win)
    seg = prediction.init_segmentor(prediction._default_cfg(14, 6), device="cuda:0", precision="tf32", state_dict=sd)
    got = agent_prediction.update_prediction(torch.from_numpy(fm).cuda(), torch.from_numpy(lm).cuda(), lmb, goal, seg, win)
    assert got.shape == ref.shape and got.dtype == np.float64
    assert float(np.abs(got - ref).max()) <= 1.5e-3                  # probabilities, tf32 path (tests/test_prednet_gpu.py)
    assert np.array_equal(got == 0, ref == 0) or float(np.abs(got - ref).max()) <= 1.5e-3
    dev = agent_prediction.update_prediction(torch.from_numpy(fm).cuda(), torch.from_numpy(lm).cuda(), lmb, goal, seg, win,
                                             as_numpy=False)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy().astype(np.float64), got)
