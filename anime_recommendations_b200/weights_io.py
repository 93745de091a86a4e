"""Saved-weights files in the Keras-2.12 HDF5 layout (SURVEY.md §5 "Checkpoint / resume").

Files the reference writes: `wandb_anime_nn.h5` (model.save, neural_network.py:220-221) and
`wandb_main_weights.h5` (ModelCheckpoint best weights, :188-196).  They are written as real HDF5 files (minih5.py:
the subset of the format h5py emits by default) with the groups, attributes, tensor names, shapes and dtypes Keras
2.12 uses:

    /                       attrs keras_version, backend, model_config (Functional-model JSON), training_config
    /model_weights          attrs layer_names, keras_version, backend      (weights-only files: these at the root)
        /<layer>            attr  weight_names
            /<layer>/<weight>:0
          user_embedding/embeddings:0 (n_users, D) f32     anime_embedding/embeddings:0 (n_anime, D) f32
          dense/kernel:0 (1,1)   dense/bias:0 (1,)
          batch_normalization/{gamma,beta,moving_mean,moving_variance}:0 (1,) each
    /optimizer_weights      attr  weight_names;  iteration:0 (int64), Adam/m/<var>:0, Adam/v/<var>:0

TensorFlow is not installable in this image, so `tf.keras.models.load_model` on these files is untested here
(tests/golden/make_tf_golden.py checks it wherever TF 2.12 exists); the files round-trip through minih5's own
reader and are read by h5py when it is present.  A path ending in `.npz` gets the same names in a NumPy archive.
"""
from __future__ import annotations

import json

import numpy as np

from . import minih5

HDF5_MAGIC = minih5.SIGNATURE
KERAS_VERSION, BACKEND = b"2.12.0", b"tensorflow"


def _layer_weights(model):
    """[(layer name, [(weight name, array)])] in Keras layer order (neural_network.py:73-101)."""
    u, a = model.names["user"], model.names["anime"]
    w = model.get_weights()
    return [("user", []), ("anime", []),
            (u, [("%s/embeddings:0" % u, w[0])]), (a, [("%s/embeddings:0" % a, w[1])]),
            (model.names["merged"], []), ("flatten", []),
            ("dense", [("dense/kernel:0", w[2].reshape(1, 1)), ("dense/bias:0", w[3].reshape(1))]),
            ("batch_normalization", [("batch_normalization/gamma:0", w[4].reshape(1)),
                                     ("batch_normalization/beta:0", w[5].reshape(1)),
                                     ("batch_normalization/moving_mean:0", w[6].reshape(1)),
                                     ("batch_normalization/moving_variance:0", w[7].reshape(1))]),
            ("activation", [])]


def _optimizer_weights(model):
    """Keras 2.12 `optimizer.variables`: iteration, then (m, v) per trainable variable in model order."""
    u, a = model.names["user"], model.names["anime"]
    hm, hv = model.head_m.cpu().numpy(), model.head_v.cpu().numpy()
    out = [("iteration:0", np.array(model.iterations, np.int64)),
           ("Adam/m/%s/embeddings:0" % u, model.mU.cpu().numpy()), ("Adam/v/%s/embeddings:0" % u, model.vU.cpu().numpy()),
           ("Adam/m/%s/embeddings:0" % a, model.mA.cpu().numpy()), ("Adam/v/%s/embeddings:0" % a, model.vA.cpu().numpy())]
    for i, n in enumerate(("dense/kernel", "dense/bias", "batch_normalization/gamma", "batch_normalization/beta")):
        shape = (1, 1) if n == "dense/kernel" else (1,)
        out.append(("Adam/m/%s:0" % n, hm[i:i + 1].reshape(shape)))
        out.append(("Adam/v/%s:0" % n, hv[i:i + 1].reshape(shape)))
    return out


def _entries(model, include_optimizer, weights_only):
    pre = "" if weights_only else "model_weights/"
    out = {}
    for layer, ws in _layer_weights(model):
        for name, arr in ws:
            out[pre + layer + "/" + name] = arr
    if include_optimizer and not weights_only:
        for name, arr in _optimizer_weights(model):
            out["optimizer_weights/" + name] = arr
    return out


def keras_model_config(model):
    """The Functional-model config Keras 2.12 stores for neural_network.py:73-101 (plus this library's own keys
    under "animerec", which Keras ignores)."""
    u, a, mg = model.names["user"], model.names["anime"], model.names["merged"]

    def inp(name):
        return dict(class_name="InputLayer", name=name, inbound_nodes=[],
                    config=dict(batch_input_shape=[None, 1], dtype="float32", sparse=False, ragged=False, name=name))

    def emb(name, n, src):
        return dict(class_name="Embedding", name=name, inbound_nodes=[[[src, 0, 0, {}]]],
                    config=dict(name=name, trainable=True, dtype="float32", batch_input_shape=[None, 1], input_dim=int(n),
                                output_dim=int(model.dim),
                                embeddings_initializer=dict(class_name="RandomUniform",
                                                            config=dict(minval=-0.05, maxval=0.05, seed=None)),
                                embeddings_regularizer=dict(class_name="L2", config=dict(l2=float(np.float32(model.l2)))),
                                activity_regularizer=None, embeddings_constraint=None, mask_zero=False, input_length=None))
    zeros, ones = dict(class_name="Zeros", config={}), dict(class_name="Ones", config={})
    layers = [inp("user"), inp("anime"), emb(u, model.n_users, "user"), emb(a, model.n_anime, "anime"),
              dict(class_name="Dot", name=mg, inbound_nodes=[[[u, 0, 0, {}], [a, 0, 0, {}]]],
                   config=dict(name=mg, trainable=True, dtype="float32", axes=2, normalize=True)),
              dict(class_name="Flatten", name="flatten", inbound_nodes=[[[mg, 0, 0, {}]]],
                   config=dict(name="flatten", trainable=True, dtype="float32", data_format="channels_last")),
              dict(class_name="Dense", name="dense", inbound_nodes=[[["flatten", 0, 0, {}]]],
                   config=dict(name="dense", trainable=True, dtype="float32", units=1, activation="linear", use_bias=True,
                               kernel_initializer=dict(class_name="HeNormal", config=dict(seed=None)),
                               bias_initializer=zeros, kernel_regularizer=None, bias_regularizer=None,
                               activity_regularizer=None, kernel_constraint=None, bias_constraint=None)),
              dict(class_name="BatchNormalization", name="batch_normalization", inbound_nodes=[[["dense", 0, 0, {}]]],
                   config=dict(name="batch_normalization", trainable=True, dtype="float32", axis=[1], momentum=0.99,
                               epsilon=0.001, center=True, scale=True, beta_initializer=zeros, gamma_initializer=ones,
                               moving_mean_initializer=zeros, moving_variance_initializer=ones, beta_regularizer=None,
                               gamma_regularizer=None, beta_constraint=None, gamma_constraint=None)),
              dict(class_name="Activation", name="activation", inbound_nodes=[[["batch_normalization", 0, 0, {}]]],
                   config=dict(name="activation", trainable=True, dtype="float32", activation="sigmoid"))]
    return dict(class_name="Functional",
                config=dict(name="model", trainable=True, layers=layers, input_layers=[["user", 0, 0], ["anime", 0, 0]],
                            output_layers=[["activation", 0, 0]]),
                animerec=dict(names=model.names, embedding_size=model.dim, n_users=model.n_users, n_anime=model.n_anime,
                              l2_reg_factor=model.l2))


def keras_training_config():
    return dict(loss="binary_crossentropy", metrics=[[dict(class_name="MeanMetricWrapper",
                                                           config=dict(name="mse", dtype="float32", fn="mean_squared_error"))]],
                weighted_metrics=None, loss_weights=None,
                optimizer_config=dict(class_name="Custom>Adam",
                                      config=dict(name="Adam", weight_decay=None, clipnorm=None, global_clipnorm=None,
                                                  clipvalue=None, use_ema=False, ema_momentum=0.99,
                                                  ema_overwrite_frequency=None, jit_compile=False, is_legacy_optimizer=False,
                                                  learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-07,
                                                  amsgrad=False)))


def save_model(model, path, include_optimizer=True, weights_only=False):
    cfg = json.dumps(keras_model_config(model))
    if str(path).endswith(".npz"):
        ent = dict(_entries(model, include_optimizer, weights_only))
        ent["__model_config__"] = np.array(cfg)
        with open(path, "wb") as fh:
            np.savez(fh, **ent)
        return
    root = minih5.Group()
    wroot = root if weights_only else root.group("model_weights")
    layers = _layer_weights(model)
    wroot.attrs.update(layer_names=[n.encode() for n, _ in layers], backend=BACKEND, keras_version=KERAS_VERSION)
    for layer, ws in layers:
        g = wroot.group(layer)
        g.attrs["weight_names"] = [n.encode() for n, _ in ws]
        for name, arr in ws:
            g.dataset(name, arr)
    if not weights_only:
        root.attrs.update(keras_version=KERAS_VERSION, backend=BACKEND, model_config=cfg.encode())
        if include_optimizer:
            root.attrs["training_config"] = json.dumps(keras_training_config()).encode()
            ow = _optimizer_weights(model)
            og = root.group("optimizer_weights")
            og.attrs["weight_names"] = [n.encode() for n, _ in ow]
            for name, arr in ow:
                og.dataset(name, arr)
    minih5.write(path, root)


def read_container(path):
    """-> (dict name -> ndarray, config dict) for an HDF5 file (this writer's, or Keras' own when its datasets are
    contiguous) or the .npz twin."""
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic == HDF5_MAGIC:
        datasets, attrs = minih5.read(path)
        out = {k.lstrip("/"): v for k, v in datasets.items()}
        mc = attrs.get("", {}).get("model_config")
        cfg = json.loads(bytes(mc).decode("utf-8")) if mc is not None else {}
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
    own = cfg.get("animerec", cfg)
    names = own.get("names", dict(user="user_embedding", anime="anime_embedding", merged="dot_product"))
    cfg = own
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
    it = ent.get("optimizer_weights/iteration:0", ent.get(o + "iteration:0"))
    if it is not None:
        dev = model.device
        model.iterations = int(it)
        model.mU.copy_(torch.from_numpy(ent[o + "m/%s/embeddings:0" % u]).to(dev))
        model.vU.copy_(torch.from_numpy(ent[o + "v/%s/embeddings:0" % u]).to(dev))
        model.mA.copy_(torch.from_numpy(ent[o + "m/%s/embeddings:0" % a]).to(dev))
        model.vA.copy_(torch.from_numpy(ent[o + "v/%s/embeddings:0" % a]).to(dev))
        hn = ("dense/kernel", "dense/bias", "batch_normalization/gamma", "batch_normalization/beta")
        model.head_m.copy_(torch.from_numpy(np.concatenate([ent[o + "m/%s:0" % n].reshape(-1) for n in hn])).to(dev))
        model.head_v.copy_(torch.from_numpy(np.concatenate([ent[o + "v/%s:0" % n].reshape(-1) for n in hn])).to(dev))
        model.lastU.fill_(model.iterations)
        model.lastA.fill_(model.iterations)
        model._t_flush = model.iterations
    return model
