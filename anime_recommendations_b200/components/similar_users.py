"""similar_users/similar_users.py entry point: same 20 string arguments, output columns
`similar_users, similarity, favorite_animes` (similar_users.py:262-314).  Top-(n+1) over ALL users, then the
query is dropped -- exactly the reference's rule, so n+1 rows come back if the query is not among them."""
from __future__ import annotations

import argparse
import logging
import os
import random
import string

import numpy as np

from .. import data, load_model, similarity
from . import _common as C

ARGS = ["anime_df", "anime_df_type", "model", "model_type", "project_name", "main_df", "main_df_type",
        "sim_user_query", "id_query_number", "max_ratings", "sim_random_user", "num_faves", "TV_only",
        "sim_users_fn", "sim_users_type", "ID_fn", "ID_type", "ID_emb_name", "anime_emb_name", "save_sim_locally"]
logger = logging.getLogger("similar_users")


def get_fave_anime(user_id, u, a, r, anime_df, num_faves, tv_only):
    """similar_users.py:203-256: the user's top-quartile-rated anime (TV only if asked), by name."""
    sel = u == user_id
    ratings, ids = r[sel], a[sel]
    if ratings.size == 0:
        return ""
    cut = np.percentile(ratings, 75)
    order = np.argsort(-ratings[ratings >= cut], kind="stable")
    top = ids[ratings >= cut][order]
    meta = anime_df.drop_duplicates("anime_id").set_index("anime_id")
    names = []
    for i in top:
        if int(i) in meta.index and (not tv_only or meta.loc[int(i), "Type"] == "TV"):
            names.append(meta.loc[int(i), "Name"])
    return str(names[:int(num_faves)])[1:-1]


def find_similar_users(user_id, n_users, W, user_to_index, user_ids):
    q = user_to_index.get(int(user_id))
    if q is None:
        raise KeyError("user id %s is not in the trained vocabulary" % user_id)
    idx, sims = similarity.find_similar_users(W, q, int(n_users))
    return [int(user_ids[i]) for i in idx], sims


def go(args, model=None):
    u, a, r = C.read_ratings(C.artifact_path(args.main_df))
    _, user_ids = data.first_appearance_codes(u)
    user_to_index = {int(v): i for i, v in enumerate(user_ids)}
    model = model or load_model(C.artifact_path(args.model))
    W = model.get_layer(args.ID_emb_name).get_weights()[0]
    anime_df = C.read_anime_df(C.artifact_path(args.anime_df))
    if C.strtobool(args.sim_random_user):
        ids, counts = np.unique(u, return_counts=True)
        user_id = int(random.choice(ids[counts < int(args.max_ratings)].tolist()))
        logger.info("Using random user ID %s", user_id)
    else:
        user_id = int(args.sim_user_query)
    ids, sims = find_similar_users(user_id, args.id_query_number, W, user_to_index, user_ids)
    import pandas as pd
    frame = pd.DataFrame({"similar_users": ids, "similarity": sims.astype(np.float32),
                          "favorite_animes": [get_fave_anime(i, u, a, r, anime_df, args.num_faves,
                                                             C.strtobool(args.TV_only)) for i in ids]})
    filename = "User_" + str(user_id).translate({ord(c): None for c in string.whitespace}) + ".csv"
    frame.to_csv(filename, index=False)
    fn = str(user_id) + ".csv"
    pd.DataFrame([user_id], columns=["User_ID"]).to_csv(fn, index=False)
    if not C.strtobool(args.save_sim_locally) and os.environ.get("ANIMEREC_KEEP_OUTPUTS") != "1":
        os.remove(filename)
        os.remove(fn)
    return frame, filename


def main(argv=None):
    C.setup_logging("similar_users")
    ap = argparse.ArgumentParser(description="Get similar users", fromfile_prefix_chars="@")
    C.add_str_args(ap, ARGS)
    return go(ap.parse_args(argv))


if __name__ == "__main__":
    main()
