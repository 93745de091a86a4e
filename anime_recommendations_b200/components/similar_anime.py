"""similar_anime/similar_anime.py entry point: same 20 string arguments and output csv columns.

The reference ranks ALL anime (np.argsort), decorates every one of them with ten pandas look-ups, filters
by Type / genre / self and keeps the first `count` (similar_anime.py:404-468).  Here the Type/genre/self
filters become a candidate bitmask, the fused GEMV + top-k kernel ranks only the candidates, and only the
`count` winners are decorated -- the same rows in the same order (ties: DESIGN.md §3)."""
from __future__ import annotations

import argparse
import logging
import os
import random

import numpy as np

from .. import data, load_model, similarity
from . import _common as C

ARGS = ["main_df_type", "anime_df_type", "sypnopsis_df_type", "model_type", "model", "project_name", "main_df",
        "sypnopses_df", "anime_df", "anime_query", "a_query_number", "random_anime", "anime_rec_genres",
        "an_spec_genres", "types", "spec_types", "a_rec_type", "save_sim_anime", "ID_emb_name", "anime_emb_name"]
COLUMNS = ["Name", "Similarity", "Genres", "Sypnopsis", "Episodes", "Japanese name", "Studios", "Premiered",
           "Score", "Type", "Source", "Rating"]
MIN_RATINGS = 400      # hard-coded in the reference (similar_anime.py:39-41), not a config knob
logger = logging.getLogger("similar_anime")


def main_df_by_anime(args):
    """similar_anime.py:25-60: users with >= 400 ratings, first-appearance anime vocabulary."""
    u, a, r = C.read_ratings(C.artifact_path(args.main_df), min_ratings=MIN_RATINGS)
    _, anime_ids = data.first_appearance_codes(a)
    return {int(v): i for i, v in enumerate(anime_ids)}, anime_ids


def resolve_anime_id(name, anime_df):
    """similar_anime.py:386-396: cleaned name, then raw name, then cleaned column."""
    translated = C.clean(name)
    for col, key in (("Name", translated), ("Name", name), ("eng_version", translated)):
        hit = anime_df[anime_df[col] == key]
        if len(hit):
            return int(hit.anime_id.values[0]), translated
    raise IndexError("anime %r not found in the anime data frame" % name)


def anime_recs(args, name, count, anime_df, model=None):
    import pandas as pd
    syn = pd.read_csv(C.artifact_path(args.sypnopses_df), usecols=["MAL_ID", "Name", "Genres", "sypnopsis"])
    model = model or load_model(C.artifact_path(args.model))
    W = model.get_layer(args.anime_emb_name).get_weights()[0]
    anime_to_index, anime_ids = main_df_by_anime(args)
    use_types = C.checked_types(args.types, logger)
    query_id, translated = resolve_anime_id(name, anime_df)
    q = anime_to_index.get(query_id)
    if q is None:
        raise KeyError("anime id %d is not in the trained vocabulary" % query_id)
    # candidate mask over the vocabulary rows: has metadata, Type allowed, genre allowed
    meta = anime_df.drop_duplicates("anime_id").set_index("anime_id")
    in_meta = np.array([int(i) in meta.index for i in anime_ids])
    mask = in_meta.copy()
    rows = meta.reindex([int(i) for i in anime_ids])
    if C.strtobool(args.spec_types):
        if use_types is None:
            return None, translated + ".csv", translated
        mask &= rows["Type"].isin(use_types).to_numpy()
    if C.strtobool(args.an_spec_genres):
        gm = C.genre_mask(rows["Genres"].to_numpy(), args.anime_rec_genres, anime_df, logger)
        if gm is None:
            return None, translated + ".csv", translated
        mask &= gm
    idx, sims = similarity.cosine_topk_query(W, q, int(count), mask=mask, exclude=q)   # any count: multi-pass above 32
    out = []
    for i, s in zip(idx, sims):
        aid = int(anime_ids[i])
        m = rows.iloc[int(i)]
        sy = syn[syn.MAL_ID == aid].sypnopsis.values
        out.append({"Name": m["Name"], "Similarity": np.float32(s), "Genres": m["Genres"],
                    "Sypnopsis": sy[0] if len(sy) else "None", "Episodes": m["Episodes"],
                    "Japanese name": m["japanese_name"], "Studios": m["Studios"], "Premiered": m["Premiered"],
                    "Score": m["Score"], "Type": m["Type"], "Source": m["Source"], "Rating": m["Rating"]})
    return pd.DataFrame(out, columns=COLUMNS), translated + ".csv", translated


def go(args, model=None):
    anime_df = C.read_anime_df(C.artifact_path(args.anime_df), sort_by_score=True)
    if C.strtobool(args.random_anime):
        name = random.choice(anime_df["Name"].unique().tolist())
        logger.info("Using %s as random input anime", name)
    else:
        name = args.anime_query
    df, fn, _ = anime_recs(args, name, int(args.a_query_number), anime_df, model)
    if df is not None:
        df.to_csv(fn, index=False)
        if not C.strtobool(args.save_sim_anime) and os.environ.get("ANIMEREC_KEEP_OUTPUTS") != "1":
            os.remove(fn)
    return df, fn


def main(argv=None):
    C.setup_logging("similar_anime")
    ap = argparse.ArgumentParser(description="Get recommendations based on similar anime", fromfile_prefix_chars="@")
    C.add_str_args(ap, ARGS)
    return go(ap.parse_args(argv))


if __name__ == "__main__":
    main()
