"""The checkpoint format of foragax_b200/recording.py on the CPU (no GPU needed): round trip of a nested pytree
(dataclass / dict / uint32 keys / scalars), the step directories and max_to_keep pruning -- the behaviour of the
reference's orbax CheckpointManager('./rec_data', max_to_keep=100) (EcoEvoJax/sim.py:643-647,699-700)."""
import os

import numpy as np
import torch


def test_checkpoint_round_trip_and_pruning(tmp_path):
    from foragax_b200 import struct
    from foragax_b200.recording import CheckpointManager

    @struct.dataclass
    class Box:
        a: torch.Tensor
        b: dict

    tree = {"rec_data": Box(a=torch.arange(12, dtype=torch.float32).reshape(3, 4),
                            b={"key": torch.tensor([[1, 2 ** 31 + 5]], dtype=torch.int64).to(torch.uint32),
                               "flag": torch.tensor([True, False])}),
            "sim_step": 7, "life": 12.5, "nothing": None}
    mgr = CheckpointManager(str(tmp_path / "rec"), max_to_keep=3)
    assert mgr.latest_step() is None
    for step in (10, 20, 30, 40, 50):
        mgr.save(step, tree)
    assert mgr.all_steps() == [30, 40, 50] and mgr.latest_step() == 50
    assert not os.path.exists(str(tmp_path / "rec" / "10"))
    got = mgr.restore()
    assert torch.equal(got["rec_data.a"], tree["rec_data"].a)
    assert got["rec_data.b.key"].dtype == torch.uint32
    assert torch.equal(got["rec_data.b.key"].view(torch.int32), tree["rec_data"].b["key"].view(torch.int32))
    assert torch.equal(got["rec_data.b.flag"], tree["rec_data"].b["flag"])
    assert got["sim_step"] == 7 and got["life"] == 12.5 and "nothing" not in got
    old = mgr.restore(30)
    assert torch.equal(old["rec_data.a"], tree["rec_data"].a)
    # a half-written checkpoint (no tree.json) is not listed
    os.makedirs(str(tmp_path / "rec" / "60"))
    assert mgr.latest_step() == 50


def test_load_state_refuses_mismatched_templates(tmp_path):
    import pytest
    from foragax_b200 import _lib as L
    from foragax_b200.recording import CheckpointManager, load_state, save_state
    mgr = CheckpointManager(str(tmp_path / "state"))
    save_state(mgr, 3, x=torch.ones(4), key=torch.zeros(2, dtype=torch.int32).view(torch.uint32), sim_step=3)
    tmpl = {"x": torch.zeros(4), "key": torch.ones(2, dtype=torch.int32).view(torch.uint32)}
    step, scalars = load_state(mgr, tmpl)
    assert step == 3 and scalars == {"sim_step": 3}
    assert torch.equal(tmpl["x"], torch.ones(4)) and int(tmpl["key"].view(torch.int32).sum()) == 0
    with pytest.raises(L.FgxError):
        load_state(mgr, {"x": torch.zeros(5), "key": tmpl["key"]})
    with pytest.raises(L.FgxError):
        load_state(mgr, {"x": torch.zeros(4), "key": tmpl["key"], "extra": torch.zeros(1)})
