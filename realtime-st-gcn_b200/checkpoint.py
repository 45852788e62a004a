"""Checkpoint ingest / export in the reference's on-disk format.

The reference saves ``{"epoch", "model_state_dict", "optimizer_state_dict", "loss"}`` with ``torch.save``
(``Processor._save_model``, processor.py:325-334) -- the state dict of the bare module, or of
``DataParallel.module`` when several GPUs were visible -- and loads it back key by key, adding a
``module.`` prefix when the model is wrapped in ``DataParallel`` (``setup``, processor.py:39-47).  The
drop-in modules of this package keep the reference's state_dict keys, so a reference ``final.pt`` loads
directly; this module handles the envelope, the optional ``module.`` prefix (in either direction) and
reports missing / unexpected keys instead of failing half-way.
"""
import collections

import torch

_PREFIX = 'module.'


def _strip(sd):
    """Keys without a DataParallel ``module.`` prefix (checkpoints written from a wrapped model)."""
    if sd and all(k.startswith(_PREFIX) for k in sd):
        return collections.OrderedDict((k[len(_PREFIX):], v) for k, v in sd.items())
    return sd


def read_checkpoint(path_or_obj, map_location='cpu'):
    """-> ``(state_dict, meta)``.  Accepts a path / file object of a reference checkpoint, the loaded dict, or a
    bare state_dict.  ``meta`` carries ``epoch`` and ``loss`` when present."""
    obj = path_or_obj
    if not isinstance(obj, dict):
        obj = torch.load(path_or_obj, map_location=map_location, weights_only=True)
    meta = {}
    if 'model_state_dict' in obj:
        meta = {k: obj[k] for k in ('epoch', 'loss') if k in obj}
        obj = obj['model_state_dict']
    if not all(torch.is_tensor(v) for v in obj.values()):
        raise ValueError("not a single-model checkpoint (the multi-stage MS-GCN format is out of scope)")
    return _strip(obj), meta


def load_checkpoint(model, path_or_obj, strict=True, map_location=None):
    """Load a reference-format checkpoint into ``model`` (a drop-in module of this package, bare or wrapped in
    ``DataParallel``).  Returns ``meta`` (+ ``missing_keys`` / ``unexpected_keys`` when ``strict=False``).

    Tensors are copied onto the model's parameters' device; shapes must match exactly (a checkpoint trained
    with another graph or class count is an error, not a silent partial load)."""
    if map_location is None:
        p = next(model.parameters(), None)
        map_location = p.device if p is not None else 'cpu'
    sd, meta = read_checkpoint(path_or_obj, map_location)
    target = model.module if isinstance(model, torch.nn.DataParallel) else model
    own = target.state_dict()
    bad = [k for k in sd if k in own and tuple(own[k].shape) != tuple(sd[k].shape)]
    if bad:
        raise ValueError("checkpoint / model shape mismatch for: " + ", ".join(
            "%s %s vs %s" % (k, tuple(sd[k].shape), tuple(own[k].shape)) for k in bad[:5]))
    res = target.load_state_dict(sd, strict=strict)
    if not strict:
        meta = dict(meta, missing_keys=list(res.missing_keys), unexpected_keys=list(res.unexpected_keys))
    return meta


def save_checkpoint(model, path, epoch=0, loss=0.0, optimizer=None):
    """Write ``model`` in the reference's format (processor.py:325-334), readable by the reference's ``setup``."""
    target = model.module if isinstance(model, torch.nn.DataParallel) else model
    torch.save({"epoch": epoch,
                "model_state_dict": collections.OrderedDict((k, v.detach().cpu()) for k, v in target.state_dict().items()),
                "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
                "loss": loss}, path)
