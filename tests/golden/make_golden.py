"""Generate tests/golden/*.json|npz by EXECUTING the reference's own functions.

Runs only in the build container (needs /root/reference, read-only).  The
reference modules cannot be imported (top-level `import tensorflow`, `import
wandb`), so each function's source is lifted out of its file with `ast`, compiled
unchanged, and run in a namespace where only the I/O loaders (wandb / Keras
`load_model`) are replaced by in-memory stubs.  No reference source is written
into this repo -- only the numeric inputs/outputs of those functions.

    python tests/golden/make_golden.py
"""
import ast
import json
import logging
import os
import random
import re
import string
import types
import unicodedata

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def lift(path, extra=None):
    """Namespace holding every top-level function of a reference file, unmodified."""
    src = open(os.path.join(REF, path), encoding="utf-8").read()
    tree = ast.parse(src)
    ns = dict(np=np, pd=pd, re=re, string=string, unicodedata=unicodedata, ast=ast,
              random=random, os=os, logger=logging.getLogger("golden"))
    ns.update(extra or {})
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
    return ns


class FakeLayer:
    def __init__(self, w):
        self.w = w

    def get_weights(self):
        return [self.w]


class FakeModel:
    def __init__(self, tables):
        self.t = tables

    def get_layer(self, name):
        return FakeLayer(self.t[name])


def golden_lrfn():
    args = types.SimpleNamespace(max_lr="5e-05", start_lr="1e-05", min_lr="1e-05",
                                 rampup_epochs="5", sustain_epochs="0", exp_decay="0.8")
    ns = lift("neural_network/neural_network.py", dict(args=args))
    vals = [ns["lrfn"](e) for e in range(20)]
    hist = pd.read_csv(os.path.join(REF, "figure_file/anime_nn_history.csv"))
    out = dict(args=vars(args), lrfn=vals, history_lr=hist["lr"].tolist(),
               history_columns=list(hist.columns[1:]), history_rows=len(hist))
    # second parameterisation exercising the sustain branch
    args2 = types.SimpleNamespace(max_lr="0.001", start_lr="0.0001", min_lr="0.00005",
                                  rampup_epochs="3", sustain_epochs="2", exp_decay="0.5")
    ns2 = lift("neural_network/neural_network.py", dict(args=args2))
    out["case2"] = dict(args=vars(args2), lrfn=[ns2["lrfn"](e) for e in range(10)])
    json.dump(out, open(os.path.join(OUT, "lrfn.json"), "w"), indent=1)


def golden_sample_perm():
    out = {}
    for n in (10, 1000, 12345):
        df = pd.DataFrame({"i": np.arange(n)})
        out[str(n)] = df.sample(frac=1, random_state=42)["i"].tolist()[:64]
    json.dump(out, open(os.path.join(OUT, "sample_perm.json"), "w"))


def make_world(seed=3, n_users=40, n_anime=60, dim=16):
    rng = np.random.RandomState(seed)
    tables = dict(anime_embedding=rng.standard_normal((n_anime, dim)).astype(np.float32),
                  user_embedding=rng.standard_normal((n_users, dim)).astype(np.float32))
    anime_ids = (np.arange(n_anime) * 7 + 5).tolist()           # MAL ids, not 0..n
    user_ids = (np.arange(n_users) * 11 + 3).tolist()
    types_ = ["TV", "OVA", "Movie", "Special", "ONA", "Music"]
    genres = ["Action, Comedy", "Drama", "Slice of Life, Comedy", "Vampire, Action", "Romance",
              "Sci-Fi, Drama"]
    anime_df = pd.DataFrame(dict(
        anime_id=anime_ids, eng_version=["anime%d" % i for i in range(n_anime)],
        Score=rng.uniform(5, 9, n_anime).round(2), Genres=[genres[i % 6] for i in range(n_anime)],
        Episodes=rng.randint(1, 50, n_anime), Premiered="Spring 2000", Studios="S",
        japanese_name=["jp%d" % i for i in range(n_anime)], Name=["Anime %d" % i for i in range(n_anime)],
        Type=[types_[(i * 5) % 6] for i in range(n_anime)], Source="Manga", Rating="PG",
        Members=rng.randint(10, 1000, n_anime)))
    syn = pd.DataFrame(dict(MAL_ID=anime_ids, Name=anime_df["Name"], Genres=anime_df["Genres"],
                            sypnopsis=["syn %d" % i for i in range(n_anime)]))
    # ratings frame: every user rates a random subset; first-appearance order == id order
    rows = []
    for ui, uid in enumerate(user_ids):
        k = rng.randint(n_anime // 2, n_anime)
        pick = np.arange(n_anime) if ui == 0 else np.sort(rng.choice(n_anime, k, replace=False))
        for a in pick:
            rows.append((uid, anime_ids[a], rng.randint(0, 11) / 10.0))
    df = pd.DataFrame(rows, columns=["user_id", "anime_id", "rating"])
    return tables, anime_df, syn, df, anime_ids, user_ids


def golden_similarity():
    tables, anime_df, syn, df, anime_ids, user_ids = make_world()
    model = FakeModel(tables)
    out = dict(anime_table=tables["anime_embedding"], user_table=tables["user_embedding"],
               anime_ids=np.array(anime_ids), user_ids=np.array(user_ids),
               anime_type=anime_df["Type"].values.astype("U"), anime_genres=anime_df["Genres"].values.astype("U"))

    # --- get_weights / extract_weights (similar_users.py:75-101, neural_network.py:128-138)
    args = types.SimpleNamespace(anime_emb_name="anime_embedding", ID_emb_name="user_embedding")
    su = lift("similar_users/similar_users.py", dict(args=args))
    aw, uw = su["get_weights"](model)
    out["anime_weights_norm"], out["user_weights_norm"] = aw, uw
    nn = lift("neural_network/neural_network.py", dict(args=args))
    out["extract_weights_user"] = nn["extract_weights"]("user_embedding", model)

    # --- find_similar_users (similar_users.py:262-314), get_fave_anime stubbed
    su["get_fave_anime"] = lambda *a, **k: ""
    user_to_index = {v: c for c, v in enumerate(user_ids)}
    index_to_user = {c: v for c, v in enumerate(user_ids)}
    su_q, su_ids, su_sims = [], [], []
    for q in (user_ids[0], user_ids[7], user_ids[-1]):
        frame, fname, _ = su["find_similar_users"](q, 5, 3, True, df, anime_df, user_to_index,
                                                   index_to_user, uw)
        su_q.append(q)
        su_ids.append(frame["similar_users"].tolist())
        su_sims.append(frame["similarity"].tolist())
    out["su_query"], out["su_ids"], out["su_sims"] = np.array(su_q), np.array(su_ids), np.array(su_sims, dtype=np.float32)

    # --- anime_recs (similar_anime.py:364-471) with loaders stubbed, type filter on
    def run_anime_recs(spec_types, types_list, an_spec_genres, genres3, name, count):
        a = types.SimpleNamespace(anime_emb_name="anime_embedding", ID_emb_name="user_embedding",
                                  types=str(types_list), spec_types=spec_types,
                                  an_spec_genres=an_spec_genres, anime_rec_genres=str(genres3))
        sa = lift("similar_anime/similar_anime.py", dict(args=a))
        sa["get_sypnopses_df"] = lambda: syn
        sa["get_model"] = lambda: model
        a2i = {v: c for c, v in enumerate(anime_ids)}
        i2a = {c: v for c, v in enumerate(anime_ids)}
        sa["main_df_by_anime"] = lambda: (df, a2i, i2a)
        frame = sa["anime_recs"](name, count, anime_df)[0]
        return frame

    f1 = run_anime_recs(True, ["TV", "Movie"], False, [None, None, None], "Anime 4", 10)
    out["sa1_names"] = f1["Name"].values.astype("U")
    out["sa1_sims"] = f1["Similarity"].values.astype(np.float32)
    out["sa1_columns"] = np.array(list(f1.columns)).astype("U")
    f2 = run_anime_recs(False, ["TV"], False, [None, None, None], "Anime 17", 7)
    out["sa2_names"] = f2["Name"].values.astype("U")
    out["sa2_sims"] = f2["Similarity"].values.astype(np.float32)
    f3 = run_anime_recs(True, ["TV", "Special", "ONA"], True, ["None", "comedy", "va#mpire"], "Anime 9", 6)
    out["sa3_names"] = f3["Name"].values.astype("U")
    out["sa3_sims"] = f3["Similarity"].values.astype(np.float32)

    # --- model_recs index plumbing (model_recs.py:132-192): get_unwatched / get_user_anime_arr
    mr = lift("model_recs/model_recs.py", dict(args=types.SimpleNamespace()))
    q = user_ids[5]
    unwatched = mr["get_unwatched"](df, anime_df, q)
    ua = mr["get_user_anime_arr"](df, anime_df, q, unwatched)
    out["mr_user"] = np.array(q)
    out["mr_unwatched_sorted"] = np.sort(np.array(unwatched).reshape(-1))
    out["mr_user_arr0"] = np.array(ua[0][0])
    out["ratings_user_id"] = df["user_id"].values
    out["ratings_anime_id"] = df["anime_id"].values
    out["ratings_rating"] = df["rating"].values
    np.savez_compressed(os.path.join(OUT, "similarity_world.npz"), **out)


def golden_preprocess():
    rng = np.random.RandomState(5)
    n = 400
    raw = pd.DataFrame(dict(user_id=rng.randint(0, 12, n), anime_id=rng.randint(0, 30, n),
                            rating=rng.randint(0, 11, n).astype(float),
                            watching_status=rng.randint(1, 7, n), watched_episodes=rng.randint(0, 5, n)))
    raw.loc[rng.choice(n, 10, replace=False), "rating"] = np.nan
    raw = pd.concat([raw, raw.iloc[:15]], ignore_index=True)
    cases = {}
    for name, kw in (("plain", dict(drop_unwatched=False, drop_plan=False, num_reviews="30")),
                     ("strict", dict(drop_unwatched=True, drop_plan=True, num_reviews="20"))):
        args = types.SimpleNamespace(**kw)
        pp = lift("preprocess/preprocess.py", dict(args=args))
        d = pp["drop_useless"](raw.copy())
        d = pp["scale_ratings"](d)
        cases[name] = dict(args=kw, out=d.to_dict(orient="list"), index=d.index.tolist())
    json.dump(dict(raw=raw.where(raw.notna(), None).to_dict(orient="list"), cases=cases),
              open(os.path.join(OUT, "preprocess.json"), "w"))


def golden_user_recs():
    """similar_user_recs / fave_genres / fave_sources / get_fave_df (user_recs.py:348-404, 708-794) executed on the
    synthetic world; metadata decoration (get_anime_frame, get_sypnopsis) runs unmodified too."""
    tables, anime_df, syn, df, anime_ids, user_ids = make_world()
    anime_df = anime_df.copy()
    anime_df["sources"] = anime_df["Source"]
    args = types.SimpleNamespace(user_recs_fn="recs.csv", ID_spec_genres=False)
    ur = lift("user_recs/user_recs.py", dict(args=args))
    out = dict(query=[], sim=[], rec_anime_id=[], rec_count=[], fav_user=[], fav_names=[])
    rng = np.random.RandomState(17)
    for q in (user_ids[2], user_ids[11], user_ids[30]):
        sims = [int(x) for x in rng.choice([u for u in user_ids if u != q], 6, replace=False)]
        sim_df = pd.DataFrame(dict(similar_users=sims, similarity=np.linspace(0.9, 0.5, 6)))
        pref = ur["get_fave_df"](ur["fave_genres"](q, df, anime_df), ur["fave_sources"](q, df, anime_df))
        frame, fname = ur["similar_user_recs"](q, sim_df, syn, df, None, None, anime_df, 12, None, None, pref)
        out["query"].append(q)
        out["sim"].append(sims)
        out["rec_anime_id"].append(frame["anime_id"].tolist())
        out["rec_count"].append(frame["n_user_prefs"].tolist())
    for u in user_ids[:8]:
        fav = ur["fave_genres"](u, df, anime_df)
        out["fav_user"].append(u)
        out["fav_names"].append(sorted(fav["eng_version"].tolist()))
    json.dump(out, open(os.path.join(OUT, "user_recs.json"), "w"))


if __name__ == "__main__":
    golden_lrfn()
    golden_sample_perm()
    golden_similarity()
    golden_preprocess()
    golden_user_recs()
    print("golden fixtures written to", OUT)
