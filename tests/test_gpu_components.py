"""Component entry points (same --arg names as the reference CLIs) end to end on the GPU: the golden world
of tests/golden/similarity_world.npz -- outputs of the reference's own anime_recs / find_similar_users run on
that world -- must come back from `components.similar_anime` / `components.similar_users`; a tiny
preprocess -> neural_network -> model_recs chain must run and agree with the oracle."""
import os
import types

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import similarity as osim
from oracle import train as ot

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200.components import (model_recs, neural_network, preprocess, similar_anime,
                                                       similar_users)


@pytest.fixture()
def world_dir(golden_dir, tmp_path, monkeypatch):
    z = np.load(os.path.join(golden_dir, "similarity_world.npz"))
    n = len(z["anime_ids"])
    anime = pd.DataFrame({"MAL_ID": z["anime_ids"], "Name": ["Anime %d" % i for i in range(n)],
                          "English name": ["anime%d" % i for i in range(n)],
                          "Japanese name": ["jp%d" % i for i in range(n)], "Score": np.linspace(9, 5, n).round(2),
                          "Genres": z["anime_genres"], "Episodes": 12, "Premiered": "Spring 2000", "Studios": "S",
                          "Type": z["anime_type"], "Source": "Manga", "Rating": "PG", "Members": 100})
    anime.to_csv(tmp_path / "all_anime.csv", index=False)
    pd.DataFrame({"MAL_ID": z["anime_ids"], "Name": anime["Name"], "Genres": anime["Genres"],
                  "sypnopsis": ["syn %d" % i for i in range(n)]}).to_csv(tmp_path / "synopses.csv", index=False)
    pd.DataFrame({"user_id": z["ratings_user_id"], "anime_id": z["ratings_anime_id"],
                  "rating": z["ratings_rating"]}).to_parquet(tmp_path / "preprocessed_stats.parquet", index=False)
    m = ar.EmbeddingDotModel(len(z["user_ids"]), n, z["anime_table"].shape[1], seed=0, dense_kernel=1.3)
    m.set_weights([z["user_table"], z["anime_table"], np.array([[1.3]]), np.zeros(1), np.ones(1), np.zeros(1),
                   np.zeros(1), np.ones(1)])
    m.save(str(tmp_path / "wandb_anime_nn.h5"))
    monkeypatch.setenv("ANIMEREC_ARTIFACT_DIR", str(tmp_path))
    monkeypatch.setenv("ANIMEREC_KEEP_OUTPUTS", "1")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(similar_anime, "MIN_RATINGS", 0)
    return z, tmp_path


def _sa_args(**kw):
    d = dict(main_df_type="t", anime_df_type="t", sypnopsis_df_type="t", model_type="h5", model="wandb_anime_nn.h5:v12",
             project_name="p", main_df="preprocessed_stats.parquet:v2", sypnopses_df="synopses.csv:v0",
             anime_df="all_anime.csv:v0", anime_query="Anime 4", a_query_number="10", random_anime="False",
             anime_rec_genres="[None, None, None]", an_spec_genres="False", types="['TV', 'Movie']", spec_types="True",
             a_rec_type="t", save_sim_anime="True", ID_emb_name="user_embedding", anime_emb_name="anime_embedding")
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.mark.parametrize("case,kw", [
    ("sa1", dict()),
    ("sa2", dict(anime_query="Anime 17", a_query_number="7", spec_types="False", types="['TV']")),
    ("sa3", dict(anime_query="Anime 9", a_query_number="6", types="['TV', 'Special', 'ONA']", an_spec_genres="True",
                 anime_rec_genres="['None', 'comedy', 'va#mpire']")),
])
def test_similar_anime_component_reproduces_reference_frames(world_dir, case, kw):
    z, tmp = world_dir
    df, fn = similar_anime.go(_sa_args(**kw))
    assert list(df.columns) == list(z["sa1_columns"])
    assert df["Name"].tolist() == z[case + "_names"].tolist()
    np.testing.assert_allclose(df["Similarity"].to_numpy(np.float32), z[case + "_sims"], rtol=0, atol=3e-6)
    back = pd.read_csv(tmp / fn)                                   # the csv the reference writes
    assert back["Name"].tolist() == df["Name"].tolist() and list(back.columns) == list(df.columns)


def test_similar_anime_invalid_genre_logs_and_returns_none(world_dir):
    df, _ = similar_anime.go(_sa_args(an_spec_genres="True", anime_rec_genres="['None', 'notagenre', 'comedy']"))
    assert df is None                                              # similar_anime.py:293-298


def test_similar_users_component_reproduces_reference_frames(world_dir):
    z, tmp = world_dir
    for q, ids, sims in zip(z["su_query"], z["su_ids"], z["su_sims"]):
        args = types.SimpleNamespace(
            anime_df="all_anime.csv:v0", anime_df_type="t", model="wandb_anime_nn.h5:v12", model_type="h5",
            project_name="p", main_df="preprocessed_stats.parquet:v2", main_df_type="t", sim_user_query=str(int(q)),
            id_query_number="5", max_ratings="600", sim_random_user="False", num_faves="3", TV_only="True",
            sim_users_fn="sim_users.csv", sim_users_type="t", ID_fn="ID.csv", ID_type="t", ID_emb_name="user_embedding",
            anime_emb_name="anime_embedding", save_sim_locally="True")
        frame, fn = similar_users.go(args)
        assert list(frame.columns) == ["similar_users", "similarity", "favorite_animes"]
        assert frame["similar_users"].tolist() == ids.tolist()
        np.testing.assert_allclose(frame["similarity"].to_numpy(np.float32), sims, rtol=0, atol=3e-6)
        assert fn == "User_%d.csv" % int(q) and os.path.exists(tmp / fn) and os.path.exists(tmp / ("%d.csv" % int(q)))


def test_model_recs_component_matches_oracle_predict(world_dir):
    z, tmp = world_dir
    user = int(z["mr_user"])
    args = types.SimpleNamespace(
        main_df="preprocessed_stats.parquet:v2", main_df_type="t", project_name="p", anime_df="all_anime.csv:v0",
        anime_df_type="t", sypnopsis_df="synopses.csv:v0", sypnopsis_df_type="t", model="wandb_anime_nn.h5:v12",
        model_type="h5", model_user_query=str(user), random_user="False", model_recs_fn="model_recs.csv",
        save_model_recs="True", model_num_recs="8", anime_types="['TV', 'Movie', 'OVA']", specify_types="True",
        model_genres="['None', 'None', 'None']", specify_genres="False", model_ID_flow="False", model_ID_conf="True",
        model_recs_type="t", flow_ID="x", flow_ID_type="t")
    frame, fn = model_recs.go(args)
    # oracle: predict every unwatched anime of the allowed types, sort by prediction
    st = ot.init_state(len(z["user_ids"]), len(z["anime_ids"]), z["anime_table"].shape[1], seed=0, w=1.3)
    st.U[:], st.A[:] = z["user_table"], z["anime_table"]
    uidx = z["user_ids"].tolist().index(user)
    unwatched = z["mr_unwatched_sorted"]
    ok = np.isin(z["anime_type"][unwatched], ["TV", "Movie", "OVA"])
    cand = unwatched[ok]
    pred = osim.model_scores(st, uidx, cand)
    order = np.argsort(-pred, kind="stable")[:8]
    assert frame["anime_id"].tolist() == z["anime_ids"][cand[order]].tolist()
    np.testing.assert_allclose(frame["Prediction"].to_numpy(np.float32), pred[order], rtol=0, atol=2e-6)
    assert fn == "User_ID_%d_model_recs.csv" % user


def test_preprocess_then_neural_network_components_train_and_save(tmp_path, monkeypatch):
    monkeypatch.setenv("ANIMEREC_ARTIFACT_DIR", str(tmp_path))
    monkeypatch.chdir(tmp_path)
    rng = np.random.RandomState(0)
    n = 6000
    raw = pd.DataFrame(dict(user_id=rng.randint(0, 40, n), anime_id=rng.randint(0, 90, n),
                            rating=rng.randint(0, 11, n).astype(float), watching_status=rng.randint(1, 7, n),
                            watched_episodes=rng.randint(0, 5, n)))
    raw.to_parquet(tmp_path / "user_stats.parquet", index=False)
    pp = types.SimpleNamespace(raw_stats="user_stats.parquet:latest", project_name="p",
                               preprocessed_stats="preprocessed_stats.parquet", preprocessed_artifact_type="t",
                               preprocessed_artifact_description="d", num_reviews="100", drop_half_watched="False",
                               save_clean_locally="False", drop_unwatched="False", drop_plan="False")
    out = preprocess.go(pp)
    clean = pd.read_parquet(out)
    assert clean["rating"].between(0, 1).all() and clean["user_id"].value_counts().min() >= 100
    argv = []
    cfg = dict(test_size="500", TPU_INIT="False", embedding_size="32", kernel_initializer="he_normal",
               activation_function="sigmoid", model_loss="binary_crossentropy", optimizer="Adam", start_lr="1e-3",
               min_lr="1e-3", max_lr="5e-3", batch_size="512", rampup_epochs="2", sustain_epochs="0", exp_decay="0.8",
               weights_artifact="wandb_main_weights.h5", save_weights_only="True", checkpoint_metric="val_loss",
               save_freq="epoch", mode="min", save_best_weights="True", verbose="0", epochs="3", save_model="True",
               model_name="./wandb_anime_nn.h5", input_data="preprocessed_stats.parquet:v2", project_name="p",
               model_artifact="wandb_anime_nn.h5", history_csv="wandb_anime_nn_history.csv",
               ID_emb_name="user_embedding", anime_emb_name="anime_embedding", merged_name="dot_product",
               main_df_type="t", model_type="h5", weights_type="h5", history_type="t", model_metrics='["mse"]',
               l2_reg_factor="0.0001", seed="3")
    for k, v in cfg.items():
        argv += ["--" + k, v]
    model, history = neural_network.go(neural_network.parse(argv))
    hist = pd.read_csv(tmp_path / "wandb_anime_nn_history.csv", index_col=0)
    assert list(hist.columns) == ["loss", "mse", "val_loss", "val_mse", "lr"] and len(hist) == 3   # anime_nn_history.csv layout
    np.testing.assert_allclose(hist["lr"].to_numpy(), [ar.lrfn(e, 1e-3, 1e-3, 5e-3, 2, 0, 0.8) for e in range(3)], rtol=1e-6)
    assert hist["loss"].iloc[-1] < hist["loss"].iloc[0]
    m2 = ar.load_model(str(tmp_path / "wandb_anime_nn.h5"))
    np.testing.assert_array_equal(m2.get_layer("user_embedding").get_weights()[0],
                                  model.get_layer("user_embedding").get_weights()[0])
    assert os.path.exists(tmp_path / "wandb_main_weights.h5") and os.path.exists(tmp_path / "history.json")
