"""oracle/agent_prediction.py against the golden output of the reference's own ``Agent_State.update_prediction`` source
(tests/golden/make_update_prediction_golden.py: the method is cut out of nav/agent/agent_state.py with ast and executed
on a stub; bit-equality with the oracle is required when the fixture is written)."""
import os

import numpy as np

from oracle import agent_prediction as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "update_prediction.npz")


def test_oracle_matches_reference_golden():
    g = np.load(GOLD)
    n = sum(1 for k in g.files if k.startswith("meta_"))
    assert n >= 4
    for i in range(n):
        full, local, win, seed, goal = (int(v) for v in g[f"meta_{i}"])
        fm, lm, lmb = O.synth_state(seed, full, local)
        tp = O.update_prediction(fm, lm, lmb, goal, O.fake_prediction, win)
        ref = g[f"target_pred_{i}"]
        assert tp.dtype == ref.dtype and np.array_equal(tp, ref)
        assert np.array_equal(fm[:, lmb[0]:lmb[1], lmb[2]:lmb[3]], lm)      # the stamp into the full map
        assert (tp[lm[1] >= 0.5] == 0).all()                                # explored cells are masked out


def test_dtype_follows_the_branch():
    fm, lm, lmb = O.synth_state(9, 64, 32)
    assert O.update_prediction(fm.copy(), lm, lmb, 1, O.fake_prediction, 48).dtype == np.float64   # crop -> float64 canvas
    assert O.update_prediction(fm.copy(), lm, lmb, 1, O.fake_prediction, 64).dtype == np.float32   # whole map -> model output
