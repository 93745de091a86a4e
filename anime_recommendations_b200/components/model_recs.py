"""model_recs/model_recs.py entry point: same 23 string arguments; predicted rating of every anime the
user has not rated, Type/genre filtered, top `model_num_recs` by Prediction (model_recs.py:132-192,373-456).

The reference calls model.predict on (user, every unwatched anime) and sorts; Prediction is a monotone map of
cos(u, a), so the candidates are ranked by the fused scoring path (tensor-core candidates, fp32 re-rank) and
only the winners go through the exact forward (similarity.score_topk)."""
from __future__ import annotations

import argparse
import logging
import os
import random

import numpy as np

from .. import data, load_model, similarity
from . import _common as C

ARGS = ["main_df", "main_df_type", "project_name", "anime_df", "anime_df_type", "sypnopsis_df", "sypnopsis_df_type",
        "model", "model_type", "model_user_query", "random_user", "model_recs_fn", "save_model_recs",
        "model_num_recs", "anime_types", "specify_types", "model_genres", "specify_genres", "model_ID_flow",
        "model_ID_conf", "model_recs_type", "flow_ID", "flow_ID_type"]
logger = logging.getLogger("model_recs")


def select_user(args, u):
    """model_recs.py:348-370."""
    if C.strtobool(args.model_ID_flow):
        import pandas as pd
        return int(pd.read_csv(C.artifact_path(args.flow_ID)).values[0][0])
    if C.strtobool(args.model_ID_conf):
        return int(args.model_user_query)
    return int(random.choice(np.unique(u).tolist()))


def go(args, model=None):
    import pandas as pd
    u, a, r = C.read_ratings(C.artifact_path(args.main_df))
    ucode, user_ids = data.first_appearance_codes(u)
    acode, anime_ids = data.first_appearance_codes(a)
    anime_df = C.read_anime_df(C.artifact_path(args.anime_df))
    syn = pd.read_csv(C.artifact_path(args.sypnopsis_df), usecols=["MAL_ID", "Name", "Genres", "sypnopsis"])
    model = model or load_model(C.artifact_path(args.model))
    user = select_user(args, u)
    uidx = {int(v): i for i, v in enumerate(user_ids)}.get(user)
    if uidx is None:
        raise KeyError("user id %s is not in the trained vocabulary" % user)
    # candidates: vocabulary rows with metadata (model_recs.py:144-155 intersects with the anime frame),
    # optionally of the requested Types / genres; the user's own ratings are the watched mask
    meta = anime_df.drop_duplicates("anime_id").set_index("anime_id")
    rows = meta.reindex([int(i) for i in anime_ids])
    mask = np.array([int(i) in meta.index for i in anime_ids])
    if C.strtobool(args.specify_types):
        mask &= rows["Type"].isin(C.literal(args.anime_types)).to_numpy()
    if C.strtobool(args.specify_genres):
        gm = C.genre_mask(rows["Genres"].to_numpy(), args.model_genres, anime_df, logger)
        if gm is not None:
            mask &= gm
    watched = np.unique(acode[ucode == uidx]).astype(np.int32)
    k = min(int(args.model_num_recs), similarity._capi.MAX_K)
    idx, pred = similarity.score_topk(model, [uidx], np.array([0, watched.size]), watched, k, cand_mask=mask)
    out = []
    for i, p in zip(idx[0], pred[0]):
        if i < 0:
            continue
        m, aid = rows.iloc[int(i)], int(anime_ids[int(i)])
        sy = syn[syn.MAL_ID == aid].sypnopsis.values
        out.append({"Name": m["Name"], "Prediction": np.float32(p), "Genres": m["Genres"], "Source": m["Source"],
                    "anime_id": aid, "Sypnopsis": sy[0] if len(sy) else "None", "Episodes": m["Episodes"],
                    "Japanese name": m["japanese_name"], "Studios": m["Studios"], "Premiered": m["Premiered"],
                    "Score": m["Score"], "Type": m["Type"]})
    frame = pd.DataFrame(out)
    filename = "User_ID_" + str(user) + "_" + args.model_recs_fn
    frame.to_csv(filename, index=False)
    if not C.strtobool(args.save_model_recs) and os.environ.get("ANIMEREC_KEEP_OUTPUTS") != "1":
        os.remove(filename)
    return frame, filename


def main(argv=None):
    C.setup_logging("model_recs")
    ap = argparse.ArgumentParser(description="Get model-based recommendations", fromfile_prefix_chars="@")
    C.add_str_args(ap, ARGS)
    return go(ap.parse_args(argv))


if __name__ == "__main__":
    main()
