"""Recording ring and checkpoints of the example drivers.

Reference: ``EcoEvoJax/sim.py:555-595`` (``Rec_Data``, ``init_rec_data``, ``record_data``: eight per-frame leaves
written with ``.at[frame].set``), ``sim.py:643-647,695-700`` (the ring is saved with an orbax
``CheckpointManager('./rec_data', max_to_keep=100)`` once ``num_rec_frames`` frames are in) and the same pattern in
``wolf_sheep_grass/sim.ipynb:301-362,406-410,530-540``.

Here the ring lives on the device and a frame is recorded by ONE native launch per agent set (``fgx_record_frame``;
the frame index is a device scalar, so recording can sit inside a captured tick).  orbax is a third-party dependency
of the reference that is not installed in this image, so checkpoints use a small self-describing on-disk format
instead -- ``<dir>/<step>/tree.json`` (the pytree structure, dtypes, shapes) + ``<dir>/<step>/leaves.npz`` -- with the
reference manager's behaviour: one directory per step, ``max_to_keep`` newest kept, ``latest_step`` / ``restore``.
The reference never checkpoints simulation STATE (it cannot resume); :func:`save_state` / :func:`load_state` add
that on the same format: every leaf of the agent sets plus the driver's key and counters.
"""
import json
import os
import shutil

import numpy as np
import torch

from . import _lib as L
from . import struct


# ---- the ring -------------------------------------------------------------------------------------------------
@struct.dataclass
class Rec_Data:  # sim.py:555-565
    forager_xs: torch.Tensor
    forager_ys: torch.Tensor
    forager_angs: torch.Tensor
    forager_rads: torch.Tensor
    resource_xs: torch.Tensor
    resource_ys: torch.Tensor
    resource_rads: torch.Tensor
    resource_blossoms: torch.Tensor


def init_rec_data(frames, forager_set, resource_set):  # sim.py:567-579
    dev = L.device()
    nf, nr = int(forager_set.num_total_agents), int(resource_set.num_total_agents)
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
    return Rec_Data(forager_xs=z(frames, nf, 1), forager_ys=z(frames, nf, 1), forager_angs=z(frames, nf, 1),
                    forager_rads=z(frames, nf, 1), resource_xs=z(frames, nr, 1), resource_ys=z(frames, nr, 1),
                    resource_rads=z(frames, nr, 1), resource_blossoms=z(frames, nr))


def record_leaves(pairs, frame):
    """``ring[frame mod frames] = leaf`` for every ``(leaf, ring)`` pair of ONE agent set (all leaves have the
    same number of rows) in one launch.  ``frame``: Python int or int32 device scalar."""
    if not pairs:
        return
    n = int(pairs[0][0].shape[0])
    frames = int(pairs[0][1].shape[0])
    arr = (L.Leaf * len(pairs))()
    keep = []
    for k, (leaf, ring) in enumerate(pairs):
        src = leaf if leaf.is_contiguous() else leaf.contiguous()
        if src.dtype != ring.dtype:
            src = src.to(ring.dtype)  # .at[].set casts the update to the ring's dtype (blossom: bool -> float32)
        keep.append(src)
        if int(src.shape[0]) != n or int(ring.shape[0]) != frames or ring.numel() != frames * src.numel():
            raise L.FgxError("record_leaves: a ring must be [frames, *leaf.shape] and all leaves share the slot axis")
        arr[k].src, arr[k].dst = L.ptr(src), L.ptr(ring)
        arr[k].row_bytes = src.element_size() * (src.numel() // max(n, 1))
    f = L.device_i32(frame)
    L.check(L.lib.fgx_record_frame(arr, len(pairs), n, frames, L.ptr(f), L.stream()))
    del keep


def record_data(foragers, resources, rec_data_obj, frame):  # sim.py:581-594 (the ring is updated in place)
    fs, rs = foragers.state.content, resources.state.content
    r = rec_data_obj
    record_leaves([(fs['X'], r.forager_xs), (fs['Y'], r.forager_ys), (fs['ang'], r.forager_angs),
                   (fs['rad'], r.forager_rads)], frame)
    record_leaves([(rs['X'], r.resource_xs), (rs['Y'], r.resource_ys), (rs['rad'], r.resource_rads),
                   (rs['blossom'].reshape(-1), r.resource_blossoms)], frame)
    return rec_data_obj


jit_record_data = record_data


# ---- checkpoints ------------------------------------------------------------------------------------------------
def _flatten(tree, prefix=""):
    """``[(path, leaf)]`` of tensors / scalars in dataclasses, dicts, lists; ``None`` leaves are skipped."""
    out = []
    if isinstance(tree, torch.Tensor) or isinstance(tree, (int, float, bool, np.ndarray)):
        out.append((prefix or "value", tree))
    elif tree is None:
        pass
    elif struct._is_struct(tree):
        import dataclasses
        for fl in dataclasses.fields(tree):
            out += _flatten(getattr(tree, fl.name), f"{prefix}.{fl.name}" if prefix else fl.name)
    elif isinstance(tree, dict):
        for k, v in tree.items():
            out += _flatten(v, f"{prefix}.{k}" if prefix else str(k))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out += _flatten(v, f"{prefix}.{i}" if prefix else str(i))
    return out


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        t = x.detach().cpu()
        return (t.view(torch.int32).numpy().view(np.uint32) if t.dtype == torch.uint32 else t.numpy()), str(x.dtype)
    return np.asarray(x), type(x).__name__


class CheckpointManager:
    """Directory-per-step checkpoints with the behaviour of the reference's orbax manager (``sim.py:643-647``:
    ``CheckpointManager('./rec_data', checkpointer, max_to_keep=100, create=True)``; ``save(step, pytree)``)."""

    def __init__(self, directory, max_to_keep=100, create=True):
        self.directory, self.max_to_keep = str(directory), max_to_keep
        if create:
            os.makedirs(self.directory, exist_ok=True)

    def all_steps(self):
        if not os.path.isdir(self.directory):
            return []
        return sorted(int(d) for d in os.listdir(self.directory)
                      if d.isdigit() and os.path.exists(os.path.join(self.directory, d, "tree.json")))

    def latest_step(self):
        steps = self.all_steps()
        return steps[-1] if steps else None

    def save(self, step, pytree):
        """Synchronises the current stream (the leaves are read back), writes ``<dir>/<step>/`` atomically (a
        temporary directory renamed into place) and prunes the oldest steps beyond ``max_to_keep``."""
        final = os.path.join(self.directory, str(int(step)))
        tmp = final + ".tmp"
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
        arrays, meta = {}, []
        for k, (path, leaf) in enumerate(_flatten(pytree)):
            a, kind = _to_numpy(leaf)
            arrays[f"leaf{k}"] = a
            meta.append({"path": path, "kind": kind, "shape": list(a.shape), "dtype": str(a.dtype)})
        np.savez(os.path.join(tmp, "leaves.npz"), **arrays)
        with open(os.path.join(tmp, "tree.json"), "w") as f:
            json.dump({"format": "foragax_b200 checkpoint v1", "step": int(step), "leaves": meta}, f, indent=1)
        shutil.rmtree(final, ignore_errors=True)
        os.rename(tmp, final)
        if self.max_to_keep:
            for old in self.all_steps()[:-self.max_to_keep]:
                shutil.rmtree(os.path.join(self.directory, str(old)), ignore_errors=True)
        return final

    def restore(self, step=None, device=None):
        """``{path: tensor-or-scalar}`` of the checkpoint at ``step`` (default: the latest one)."""
        step = self.latest_step() if step is None else int(step)
        if step is None:
            raise FileNotFoundError(f"no checkpoint under {self.directory}")
        d = os.path.join(self.directory, str(step))
        meta = json.load(open(os.path.join(d, "tree.json")))["leaves"]
        out = {}
        with np.load(os.path.join(d, "leaves.npz")) as z:
            for k, m in enumerate(meta):
                a = z[f"leaf{k}"]
                if m["kind"].startswith("torch."):
                    if m["kind"] == "torch.uint32":
                        t = torch.from_numpy(a.view(np.int32).copy()).view(torch.uint32)
                    else:
                        t = torch.from_numpy(a.copy())
                    out[m["path"]] = t.to(device) if device is not None else t
                else:
                    out[m["path"]] = a.item() if a.shape == () else a
        return out


def save_state(manager, step, **trees):
    """Checkpoint of simulation STATE for resume: ``save_state(mgr, step, foragers=..., resources=..., key=..., ...)``."""
    return manager.save(step, trees)


def load_state(manager, template, step=None):
    """Copies a :func:`save_state` checkpoint back INTO the tensors of ``template`` (a dict of pytrees with the same
    structure, e.g. freshly created agent sets) and returns ``(step, scalars)`` -- the non-tensor entries."""
    step = manager.latest_step() if step is None else int(step)
    flat = manager.restore(step)
    scalars = {}
    have = dict(_flatten(template))
    for path, val in flat.items():
        dst = have.get(path)
        if isinstance(dst, torch.Tensor) and isinstance(val, torch.Tensor):
            if tuple(dst.shape) != tuple(val.shape) or dst.dtype != val.dtype:
                raise L.FgxError(f"checkpoint leaf {path}: {tuple(val.shape)} {val.dtype} does not fit the template")
            dst.copy_(val.to(dst.device))
        else:
            scalars[path] = val
    missing = [p for p, v in have.items() if isinstance(v, torch.Tensor) and p not in flat]
    if missing:
        raise L.FgxError(f"checkpoint at step {step} lacks {missing[:3]}...")
    return step, scalars
