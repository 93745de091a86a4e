"""Shared plumbing of the component entry points.

The reference's components fetch every input through Weights & Biases (`run.use_artifact(name, type)`);
there is no W&B server here, so artifact names ("preprocessed_stats.parquet:v2", "wandb_anime_nn.h5:v12")
resolve to local files.  Argument NAMES, string typing (`strtobool` / `ast.literal_eval`) and output file
names stay those of the reference so that `main.py`'s parameter dicts drive these entry points unchanged.
"""
from __future__ import annotations

import ast
import logging
import os
import re
import string
import unicodedata

import numpy as np

ARTIFACT_ENV = "ANIMEREC_ARTIFACT_DIR"


def strtobool(v):
    """distutils.util.strtobool as the reference's `type=lambda x: bool(strtobool(x))` uses it."""
    if isinstance(v, bool):
        return v
    s = str(v).strip().lower()
    if s in ("y", "yes", "t", "true", "on", "1"):
        return True
    if s in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError("invalid truth value %r" % (v,))


def literal(v):
    """ast.literal_eval of a string argument (lists of genres/types/metrics)."""
    return ast.literal_eval(v) if isinstance(v, str) else v


def setup_logging(name):
    logging.basicConfig(filename="./%s.log" % name, level=logging.INFO, filemode="a",
                        format="%(asctime)s-%(name)s - %(levelname)s - %(message)s",
                        datefmt="%d %b %Y %H:%M:%S %Z", force=True)
    return logging.getLogger(name)


def artifact_path(name, must_exist=True):
    """'file.ext:vN' (a W&B artifact reference) or a plain path -> local path.  Search order: the path as
    given, $ANIMEREC_ARTIFACT_DIR, ./artifacts, ./data."""
    base = str(name)
    if re.search(r":v\d+$|:latest$", base):
        base = base.rsplit(":", 1)[0]
    cands = [base]
    for d in (os.environ.get(ARTIFACT_ENV), "artifacts", "data"):
        if d:
            cands.append(os.path.join(d, os.path.basename(base)))
    for c in cands:
        if os.path.exists(c):
            return c
    if must_exist:
        raise FileNotFoundError("artifact %r not found locally (looked in %s); set %s" % (name, cands, ARTIFACT_ENV))
    return cands[1] if len(cands) > 1 and os.environ.get(ARTIFACT_ENV) else base


def add_str_args(parser, names, required=True):
    """Every reference argument is `type=str, required=True`; extras of a sibling component are tolerated."""
    for n in names:
        parser.add_argument("--" + n, type=str, required=required, help=n)


# ---------------------------------------------------------------------------------- metadata helpers
IRREGULAR = ["★", "♥", "☆", "♡", "½", "ß", "²"]


def clean(item):
    """Name normalisation of similar_anime.py:243-276: strip whitespace, punctuation, accents; lower-case."""
    if isinstance(item, list):
        return [clean(x) for x in item]
    x = str(item)
    for irr in IRREGULAR:
        x = x.replace(irr, " ")
    x = x.translate({ord(c): None for c in string.whitespace})
    x = re.sub(r"\W+", "", x)
    x = "".join(c for c in unicodedata.normalize("NFKD", x) if not unicodedata.combining(c))
    return x.lower()


def read_anime_df(path, sort_by_score=False):
    """get_anime_df of similar_anime.py:63-92 / model_recs.py:91-116 (pandas; metadata only)."""
    import pandas as pd
    df = pd.read_csv(path).replace("Unknown", np.nan)
    df["anime_id"] = df["MAL_ID"]
    df["japanese_name"] = df["Japanese name"]
    df["eng_version"] = df["Name"].map(lambda n: clean(n).lower())
    if sort_by_score:
        df = df.sort_values(by=["Score"], ascending=False, kind="quicksort", na_position="last")
    keep = ["anime_id", "eng_version", "Score", "Genres", "Episodes", "Premiered", "Studios", "japanese_name",
            "Name", "Type", "Source", "Rating", "Members"]
    return df[keep]


def all_genres(anime_df):
    """get_genres of similar_anime.py:173-191."""
    genres = anime_df["Genres"].unique().tolist()
    poss = sorted(set(re.sub(r"[\W_]", "", e) for e in set(str(genres).split())))
    rem = ["Slice", "of", "Life", "Martial", "Arts", "Super", "Power", "nan"]
    return sorted(i for i in poss + ["Slice of Life", "Super Power", "Martial Arts", "None"] if i not in rem)


def genre_mask(genres_column, use_genres_arg, anime_df_for_vocab, logger=None):
    """Row filter equivalent to by_genre (similar_anime.py:279-340, model_recs.py:270-330): keep a row when
    any of the three requested (cleaned) genres is a substring of its lower-cased, space-free Genres string;
    "none" never matches.  Returns None (after logging) when a requested genre is not a known genre."""
    use = clean(literal(use_genres_arg))
    known = clean(all_genres(anime_df_for_vocab))
    for g in use:
        if g not in known:
            if logger:
                logger.info("An invalid genre was input. Select genres from %s", known)
            return None
    col = [str(x).lower().replace(" ", "") for x in genres_column]
    return np.array([any(g != "none" and g in s for g in use) for s in col], dtype=bool)


ANIME_TYPES = ["TV", "OVA", "Movie", "Special", "ONA", "Music"]


def checked_types(types_arg, logger=None):
    """get_types of similar_anime.py:343-358."""
    use = literal(types_arg)
    for t in use:
        if t not in ANIME_TYPES:
            if logger:
                logger.info("An invalid type was input. Select from %s", ANIME_TYPES)
            return None
    return use


def read_ratings(path, min_ratings=None):
    """The preprocessed ratings frame as NumPy columns (+ the optional `>= 400 ratings` re-filter of
    similar_anime.py:39-41)."""
    import pandas as pd
    df = pd.read_parquet(path, columns=["user_id", "anime_id", "rating"])
    if min_ratings is not None:
        n = df["user_id"].value_counts(dropna=True)
        df = df[df["user_id"].isin(n[n >= int(min_ratings)].index)]
    return df["user_id"].to_numpy(), df["anime_id"].to_numpy(), df["rating"].to_numpy()
