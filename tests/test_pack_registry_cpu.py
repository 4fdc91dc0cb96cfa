"""CPU suite — host logic of the persistent packed conv operands (ops.pack_weights / ops.repack_all): which events
make a packed buffer stale, that a fresh one launches nothing, that the data-gradient operand is registered with the
forward one, and that the batched refresh stamps every entry.  The kernels themselves are stubbed (no CUDA here);
tests/test_gpu_model.py::test_batched_repack_tracks_the_weights checks the bytes on the GPU."""
import types

import pytest
import torch


@pytest.fixture()
def stub(b3d, monkeypatch):
    ops = b3d.ops
    calls = []
    monkeypatch.setattr(ops, "_call", lambda name, *a: calls.append((name, a)))
    table_builds = []

    def fake_table(flat):
        if flat.pack_table is None or flat.pack_table[3] != ops._PREC_EPOCH["n"]:
            entries = list(flat.packs.values())
            table_builds.append(len(entries))
            flat.pack_table = (torch.zeros(1, dtype=torch.int64), len(entries), 7, ops._PREC_EPOCH["n"], entries)
        return flat.pack_table

    monkeypatch.setattr(ops, "ensure_pack_table", fake_table)
    flat = types.SimpleNamespace(packs={}, pack_table=None, epoch=0, theta=torch.zeros(4))
    return ops, calls, flat, table_builds


def _weight(flat, *shape):
    w = torch.randn(*shape)
    w._b3d_flat = flat
    return w


def test_fresh_pack_launches_nothing_and_stale_events_repack(stub):
    ops, calls, flat, _ = stub
    w = _weight(flat, 3, 3, 3, 16, 16)
    buf = ops.pack_weights(w, False, 1, False)
    assert [c[0] for c in calls] == ["b3d_conv3d_pack_weights"]
    assert buf.numel() == ops.lib.b3d_conv3d_packed_elems(3, 1, 16, 16)
    # the data-gradient operand of the layer is registered together with the forward one (not packed yet)
    assert set(k[1] for k in flat.packs) == {False, True} and len(flat.packs) == 2
    assert ops.pack_weights(w, False, 1, False) is buf and len(calls) == 1            # fresh: no launch
    w.add_(1.0)                                                                        # torch-side change of the layer
    assert ops.pack_weights(w, False, 1, False) is buf and len(calls) == 2
    flat.theta.mul_(2.0)                                                               # ... of the flat buffer
    ops.pack_weights(w, False, 1, False)
    assert len(calls) == 3
    flat.epoch += 1                                                                    # optimiser kernel (per-var path)
    ops.pack_weights(w, False, 1, False)
    assert len(calls) == 4
    prev = ops.set_kd_fold(True)                                                       # layout switch (and back)
    ops.set_kd_fold(prev)
    ops.pack_weights(w, False, 1, False)
    assert len(calls) == 5
    ops.pack_weights(w, True, 1, False)                                                # dgrad operand: first use packs
    assert len(calls) == 6 and calls[-1][1][-1] == 1 and len(flat.packs) == 2


def test_repack_all_is_one_launch_and_stamps_every_entry(stub):
    ops, calls, flat, table_builds = stub
    ws = [_weight(flat, 3, 3, 3, 16, 16), _weight(flat, 1, 1, 1, 32, 16), _weight(flat, 3, 3, 3, 16, 32)]
    for w in ws:
        ops.pack_weights(w, False, 2 if w is ws[2] else 1, False)
    n_single = len(calls)
    assert n_single == 3 and len(flat.packs) == 6
    e0 = flat.epoch
    ops.repack_all(flat)                                   # what ScheduledOptim.apply_flat does after the Adam kernel
    assert flat.epoch == e0 + 1 and table_builds == [6]
    assert [c[0] for c in calls[n_single:]] == ["b3d_conv3d_pack_many"] and calls[-1][1][1:] == (6, 7)
    for w in ws:                                           # forward and data-gradient operands are all fresh now
        ops.pack_weights(w, False, 2 if w is ws[2] else 1, False)
        ops.pack_weights(w, True, 2 if w is ws[2] else 1, False)
    assert len(calls) == n_single + 1
    ops.repack_all(flat)                                   # the table is reused while the set of layers is unchanged
    assert table_builds == [6] and len(calls) == n_single + 2
    ops.pack_weights(_weight(flat, 3, 3, 3, 32, 32), False, 1, False)      # a new layer invalidates the table
    ops.repack_all(flat)
    assert table_builds == [6, 8]


def test_weights_outside_a_model_are_packed_per_call(stub):
    ops, calls, flat, _ = stub
    w = torch.randn(3, 3, 3, 16, 16)
    a, b = ops.pack_weights(w, False, 1, False), ops.pack_weights(w, False, 1, False)
    assert a is not b and len(calls) == 2 and not flat.packs
