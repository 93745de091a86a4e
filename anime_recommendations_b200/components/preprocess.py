"""preprocess/preprocess.py entry point: same 10 string arguments, local files instead of W&B artifacts.
Semantics of drop_useless + scale_ratings (preprocess.py:13-40,108-117) are in ..data; the optional
drop_half_watched filter (preprocess.py:52-105) is restated here with vectorised NumPy."""
from __future__ import annotations

import argparse
import logging
import os

import numpy as np

from .. import data
from . import _common as C

ARGS = ["raw_stats", "project_name", "preprocessed_stats", "preprocessed_artifact_type",
        "preprocessed_artifact_description", "num_reviews", "drop_half_watched", "save_clean_locally",
        "drop_unwatched", "drop_plan"]
logger = logging.getLogger("preprocess")


def drop_half_watched(cols):
    """preprocess.py:52-105: keep a row when watched_episodes >= half of the anime's maximum watched_episodes
    (anime whose maximum is 1 keep the full threshold 1)."""
    aid = np.asarray(cols["anime_id"])
    eps = np.asarray(cols["watched_episodes"], np.float64)
    uniq, inv = np.unique(aid, return_inverse=True)
    mx = np.full(uniq.size, -np.inf)
    np.maximum.at(mx, inv, eps)
    half = np.where(mx == 1, mx, mx * 0.5)[inv]
    keep = eps >= half
    return {k: np.asarray(v)[keep] for k, v in cols.items()}


def go(args):
    import pandas as pd
    df = pd.read_parquet(C.artifact_path(args.raw_stats))
    cols = {c: df[c].to_numpy() for c in data.RAW_COLUMNS}
    from .. import data_gpu            # the filters run on the GPU (bit-equal to data.py, the NumPy statement of the same rules)
    idx = data_gpu.drop_useless(cols, int(args.num_reviews), C.strtobool(args.drop_unwatched),
                                C.strtobool(args.drop_plan)).cpu().numpy()
    cols = {c: np.asarray(v)[idx] for c, v in cols.items()}
    logger.info("Useless data dropped!")
    if C.strtobool(args.drop_half_watched):
        logger.info("Dropping samples with too few episodes watched")
        cols = drop_half_watched(cols)
    cols["rating"] = data.scale_ratings(cols["rating"])
    out = pd.DataFrame({c: cols[c] for c in data.RAW_COLUMNS})
    logger.info("Final df shape is %s", out.shape)
    path = C.artifact_path(args.preprocessed_stats, must_exist=False)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    out.to_parquet(path, index=False)
    return path


def main(argv=None):
    ap = argparse.ArgumentParser(description="Preprocess a dataset", fromfile_prefix_chars="@")
    C.add_str_args(ap, ARGS)
    return go(ap.parse_args(argv))


if __name__ == "__main__":
    main()
