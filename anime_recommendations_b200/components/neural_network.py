"""neural_network/neural_network.py entry point: the same 37 string arguments (neural_network.py:298-555),
the same files out (model, best weights, history json/csv); training runs on libanimerec.so.

Differences forced by the environment, all outside the arithmetic: inputs come from local files instead of
W&B artifacts; no loss plot is uploaded.  `--TPU_INIT True` is the reference's distribution-strategy switch
(neural_network.py:142-147,173-182): here it selects multi-GPU data parallelism and must be launched with one
process per GPU (`python -m torch.distributed.run --nproc-per-node N -m
anime_recommendations_b200.components.neural_network ...`); as in the reference the global batch is
batch_size x replicas and max_lr is scaled by the replica count.  Without torchrun it is an error, not a silent
single-GPU run."""
from __future__ import annotations

import argparse
import json
import logging
import os

import numpy as np

from .. import EarlyStopping, EmbeddingDotModel, LearningRateScheduler, ModelCheckpoint, data, lrfn
from . import _common as C

ARGS = ["test_size", "TPU_INIT", "embedding_size", "kernel_initializer", "activation_function", "model_loss",
        "optimizer", "start_lr", "min_lr", "max_lr", "batch_size", "rampup_epochs", "sustain_epochs", "exp_decay",
        "weights_artifact", "save_weights_only", "checkpoint_metric", "save_freq", "mode", "save_best_weights",
        "verbose", "epochs", "save_model", "model_name", "input_data", "project_name", "model_artifact",
        "history_csv", "ID_emb_name", "anime_emb_name", "merged_name", "main_df_type", "model_type", "weights_type",
        "history_type", "model_metrics", "l2_reg_factor"]
EXTRA = dict(adam_mode="replay", seed="", shuffle="numpy")   # knobs of this implementation, all optional
logger = logging.getLogger("neural_network")


def get_df(args):
    """neural_network.py:25-63 -> data.EncodedRatings (first-appearance vocabulary, sample(random_state=42))."""
    u, a, r = C.read_ratings(C.artifact_path(args.input_data))
    from .. import data_gpu
    enc = data_gpu.encode_ratings(u, a, r)        # vocabulary + seeded shuffle on the device, same result as data.py
    logger.info("Final df shape is %s", (len(enc.user), 3))
    return enc


def init_strategy(args):
    """neural_network.py:142-147: -> number of replicas (1 without --TPU_INIT)."""
    if not C.strtobool(args.TPU_INIT):
        return 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
            raise RuntimeError("--TPU_INIT True selects multi-GPU training: launch one process per GPU with "
                               "`python -m torch.distributed.run --nproc-per-node N ...`")
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_world_size()


def build_model(args, n_users, n_anime, replicas=1):
    """neural_network.py:66-106.  Only the configuration the reference ships is implemented in CUDA."""
    if args.model_loss != "binary_crossentropy" or args.activation_function != "sigmoid" or \
            str(args.optimizer).lower() != "adam":
        raise ValueError("libanimerec implements loss=binary_crossentropy, activation=sigmoid, optimizer=Adam "
                         "(config.yaml:65-69); got %s/%s/%s" % (args.model_loss, args.activation_function, args.optimizer))
    metrics = C.literal(args.model_metrics)
    if list(metrics) != ["mse"]:
        raise ValueError("only model_metrics ['mse'] (config.yaml:88) is implemented")
    seed = int(args.seed) if str(getattr(args, "seed", "")).strip() else None
    if replicas > 1:
        from ..dist_fit import DistributedEmbeddingDotModel
        if seed is None:
            seed = 0          # every rank must draw the same initial weights
        return DistributedEmbeddingDotModel(n_users, n_anime, int(args.embedding_size),
                                            l2_reg_factor=float(args.l2_reg_factor),
                                            kernel_initializer=args.kernel_initializer, ID_emb_name=args.ID_emb_name,
                                            anime_emb_name=args.anime_emb_name, merged_name=args.merged_name, seed=seed,
                                            adam_mode=getattr(args, "adam_mode", "replay"))
    return EmbeddingDotModel(n_users, n_anime, int(args.embedding_size), l2_reg_factor=float(args.l2_reg_factor),
                             kernel_initializer=args.kernel_initializer, ID_emb_name=args.ID_emb_name,
                             anime_emb_name=args.anime_emb_name, merged_name=args.merged_name, seed=seed,
                             adam_mode=getattr(args, "adam_mode", "replay"))


def go(args):
    enc = get_df(args)
    logger.info("Data frame loaded")
    (x_train, y_train), (x_test, y_test) = data.train_test_split_tail(enc, int(args.test_size))
    replicas = init_strategy(args)
    rank0 = True
    if replicas > 1:
        import torch.distributed as dist
        rank0 = dist.get_rank() == 0
    model = build_model(args, enc.n_users, enc.n_anime, replicas)
    max_lr = float(args.max_lr) * replicas                      # neural_network.py:177
    sched = LearningRateScheduler(lambda epoch: lrfn(epoch, args.start_lr, args.min_lr, max_lr,
                                                     args.rampup_epochs, args.sustain_epochs, args.exp_decay), verbose=0)
    ckpt = ModelCheckpoint(filepath=args.weights_artifact, save_weights_only=C.strtobool(args.save_weights_only),
                           monitor=args.checkpoint_metric, save_freq=args.save_freq, mode=args.mode,
                           save_best_only=C.strtobool(args.save_best_weights), verbose=int(args.verbose))
    stop = EarlyStopping(patience=3, monitor=args.checkpoint_metric, mode=args.mode, restore_best_weights=True)
    shuffle = getattr(args, "shuffle", "numpy")
    history = model.fit(x=x_train, y=y_train, batch_size=int(args.batch_size), epochs=int(args.epochs),
                        verbose=int(args.verbose), validation_data=(x_test, y_test), callbacks=[ckpt, sched, stop],
                        shuffle=False if shuffle in ("False", "false", "none") else shuffle)
    if C.strtobool(args.save_model):
        model.save(args.model_name)
    logger.info("model trained and saved!")
    hist = history.history
    if not rank0:                                              # the files are written once, by rank 0
        return model, history
    with open("history.json", "w") as f:                      # DataFrame.to_json layout: {column: {row: value}}
        json.dump({k: {str(i): v for i, v in enumerate(vals)} for k, vals in hist.items()}, f)
    cols = list(hist.keys())
    with open(args.history_csv, "w") as f:                     # DataFrame.to_csv layout (leading index column)
        f.write("," + ",".join(cols) + "\n")
        for i in range(len(hist[cols[0]])):
            f.write(str(i) + "," + ",".join(repr(float(hist[c][i])) for c in cols) + "\n")
    if os.path.abspath(args.model_artifact) != os.path.abspath(args.model_name) and C.strtobool(args.save_model):
        model.save(args.model_artifact)
    return model, history


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Train an anime recommendation neural network", fromfile_prefix_chars="@")
    C.add_str_args(ap, ARGS)
    for k, v in EXTRA.items():
        ap.add_argument("--" + k, type=str, default=v)
    return ap.parse_args(argv)


def main(argv=None):
    C.setup_logging("neural_network")
    return go(parse(argv))


if __name__ == "__main__":
    main()
