"""Saved-weights container in the Keras-2.12 H5 name layout (SURVEY.md §5 "Checkpoint / resume").

Files the reference writes: `wandb_anime_nn.h5` (model.save, neural_network.py:220-221) and
`wandb_main_weights.h5` (ModelCheckpoint best weights, :188-196).  Tensor names, shapes and dtypes
follow Keras 2.12:

    model_weights/<layer>/<layer>/<weight>:0     (full model; weights-only files drop the prefix)
      user_embedding/embeddings:0   (n_users, D) f32      anime_embedding/embeddings:0 (n_anime, D) f32
      dense/kernel:0 (1,1)  dense/bias:0 (1,)
      batch_normalization/{gamma,beta,moving_mean,moving_variance}:0 (1,) each
    optimizer_weights/Adam/{iteration,<var>/m,<var>/v}:0

h5py/libhdf5 are not part of this image, so the container is written as HDF5 only when h5py is
importable; otherwise the SAME names/shapes/dtypes go into a NumPy .npz archive stored at the
requested path (whatever its suffix).  `load_model` sniffs the format, so downstream components
consume either unchanged.
"""
from __future__ import annotations

import json

import numpy as np

try:  # pragma: no cover - absent in the build image
    import h5py
except Exception:  # noqa: BLE001
    h5py = None

HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def _entries(model, include_optimizer, weights_only):
    u, a = model.names["user"], model.names["anime"]
    w = model.get_weights()
    pre = "" if weights_only else "model_weights/"
    out = {
        pre + "%s/%s/embeddings:0" % (u, u): w[0],
        pre + "%s/%s/embeddings:0" % (a, a): w[1],
        pre + "dense/dense/kernel:0": w[2].reshape(1, 1),
        pre + "dense/dense/bias:0": w[3].reshape(1),
        pre + "batch_normalization/batch_normalization/gamma:0": w[4].reshape(1),
        pre + "batch_normalization/batch_normalization/beta:0": w[5].reshape(1),
        pre + "batch_normalization/batch_normalization/moving_mean:0": w[6].reshape(1),
        pre + "batch_normalization/batch_normalization/moving_variance:0": w[7].reshape(1),
    }
    if include_optimizer and not weights_only:
        o = "optimizer_weights/Adam/"
        hm, hv = model.head_m.cpu().numpy(), model.head_v.cpu().numpy()
        out[o + "iteration:0"] = np.array(model.iterations, np.int64)
        out[o + "m/%s/embeddings:0" % u] = model.mU.cpu().numpy()
        out[o + "v/%s/embeddings:0" % u] = model.vU.cpu().numpy()
        out[o + "m/%s/embeddings:0" % a] = model.mA.cpu().numpy()
        out[o + "v/%s/embeddings:0" % a] = model.vA.cpu().numpy()
        for i, n in enumerate(("dense/kernel", "dense/bias", "batch_normalization/gamma", "batch_normalization/beta")):
            out[o + "m/%s:0" % n] = hm[i:i + 1]
            out[o + "v/%s:0" % n] = hv[i:i + 1]
    return out


def _config(model):
    return dict(class_name="Functional", keras_version="2.12.0", backend="tensorflow",
                layers=[model.names["user"], model.names["anime"], model.names["merged"], "flatten", "dense",
                        "batch_normalization", "activation"],
                embedding_size=model.dim, n_users=model.n_users, n_anime=model.n_anime,
                l2_reg_factor=model.l2, names=model.names,
                training_config=dict(loss="binary_crossentropy", metrics=["mse"], optimizer="Adam"))


def save_model(model, path, include_optimizer=True, weights_only=False):
    ent = _entries(model, include_optimizer, weights_only)
    cfg = json.dumps(_config(model))
    if h5py is not None:
        with h5py.File(path, "w") as f:
            for k, v in ent.items():
                f.create_dataset(k, data=v)
            f.attrs["model_config"] = cfg
            f.attrs["keras_version"] = "2.12.0"
            f.attrs["backend"] = "tensorflow"
        return
    ent = dict(ent)
    ent["__model_config__"] = np.array(cfg)
    with open(path, "wb") as fh:
        np.savez(fh, **ent)


def read_container(path):
    """-> (dict name -> ndarray, config dict) for either container flavour."""
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic == HDF5_MAGIC:
        if h5py is None:
            raise RuntimeError("%s is an HDF5 file but h5py is not installed in this environment" % path)
        out = {}
        with h5py.File(path, "r") as f:
            f.visititems(lambda n, o: out.__setitem__(n, np.array(o)) if isinstance(o, h5py.Dataset) else None)
            cfg = json.loads(f.attrs.get("model_config", "{}"))
        return out, cfg
    z = np.load(path, allow_pickle=False)
    out = {k: z[k] for k in z.files if k != "__model_config__"}
    cfg = json.loads(str(z["__model_config__"])) if "__model_config__" in z.files else {}
    return out, cfg


def _find(ent, suffix):
    hits = [k for k in ent if k.endswith(suffix)]
    hits = [k for k in hits if not k.startswith("optimizer_weights/")] or hits
    if not hits:
        raise KeyError("no tensor ending in %r in the container" % suffix)
    return ent[hits[0]]


def load_model(path, device=None, adam_mode="replay"):
    from .model import EmbeddingDotModel
    ent, cfg = read_container(path)
    names = cfg.get("names", dict(user="user_embedding", anime="anime_embedding", merged="dot_product"))
    U = _find(ent, "%s/embeddings:0" % names["user"])
    A = _find(ent, "%s/embeddings:0" % names["anime"])
    m = EmbeddingDotModel(U.shape[0], A.shape[0], U.shape[1], l2_reg_factor=cfg.get("l2_reg_factor", 1e-4),
                          ID_emb_name=names["user"], anime_emb_name=names["anime"],
                          merged_name=names.get("merged", "dot_product"), seed=0, device=device,
                          adam_mode=adam_mode, dense_kernel=1.0)
    load_into(m, path, _pre=(ent, cfg))
    return m


def load_into(model, path, _pre=None):
    import torch
    ent, cfg = _pre or read_container(path)
    u, a = model.names["user"], model.names["anime"]
    w = [_find(ent, "%s/embeddings:0" % u), _find(ent, "%s/embeddings:0" % a), _find(ent, "dense/kernel:0"),
         _find(ent, "dense/bias:0"), _find(ent, "batch_normalization/gamma:0"),
         _find(ent, "batch_normalization/beta:0"), _find(ent, "batch_normalization/moving_mean:0"),
         _find(ent, "batch_normalization/moving_variance:0")]
    if w[0].shape != (model.n_users, model.dim) or w[1].shape != (model.n_anime, model.dim):
        raise ValueError("container tables %s/%s do not match the model (%d,%d,%d)" % (
            w[0].shape, w[1].shape, model.n_users, model.n_anime, model.dim))
    model.set_weights(w)
    o = "optimizer_weights/Adam/"
    if o + "iteration:0" in ent:
        dev = model.device
        model.iterations = int(ent[o + "iteration:0"])
        model.mU.copy_(torch.from_numpy(ent[o + "m/%s/embeddings:0" % u]).to(dev))
        model.vU.copy_(torch.from_numpy(ent[o + "v/%s/embeddings:0" % u]).to(dev))
        model.mA.copy_(torch.from_numpy(ent[o + "m/%s/embeddings:0" % a]).to(dev))
        model.vA.copy_(torch.from_numpy(ent[o + "v/%s/embeddings:0" % a]).to(dev))
        hn = ("dense/kernel", "dense/bias", "batch_normalization/gamma", "batch_normalization/beta")
        model.head_m.copy_(torch.from_numpy(np.concatenate([ent[o + "m/%s:0" % n] for n in hn])).to(dev))
        model.head_v.copy_(torch.from_numpy(np.concatenate([ent[o + "v/%s:0" % n] for n in hn])).to(dev))
        model.lastU.fill_(model.iterations)
        model.lastA.fill_(model.iterations)
        model._t_flush = model.iterations
    return model
