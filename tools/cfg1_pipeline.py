"""BASELINE configs[0] end to end on the GPU path (functional run + parity spot checks):

    synthetic user_stats.parquet (the reference's schema; the real file is absent from the checkout,
    SURVEY F3) -> components.preprocess -> components.neural_network (1 epoch, embedding 128, batch 10000,
    config.yaml defaults) -> components.similar_anime top-10 for one anime.

    python tools/cfg1_pipeline.py [n_ratings=7000000] [workdir=/tmp/cfg1]

Prints one JSON line with the stage wall times and the parity checks made along the way."""
import json
import os
import sys
import time
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_raw(n, n_users=25_000, n_anime=17_560, seed=0):
    """Zipf-ish user activity (so that the num_reviews=400 filter bites), every anime present, ratings 0..10."""
    rng = np.random.RandomState(seed)
    w = 1.0 / np.arange(1, n_users + 1) ** 0.8
    users = rng.choice(n_users, n, p=w / w.sum())
    anime = np.r_[np.arange(n_anime), rng.randint(0, n_anime, n - n_anime)]
    order = np.argsort(users, kind="stable")                      # the real file is sorted by user id
    return pd.DataFrame(dict(user_id=users[order] + 1, anime_id=(anime[order] * 3 + 1),
                             rating=rng.randint(0, 11, n)[order].astype(np.float64),
                             watching_status=rng.randint(1, 7, n)[order], watched_episodes=rng.randint(0, 30, n)[order]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 7_000_000
    work = sys.argv[2] if len(sys.argv) > 2 else "/tmp/cfg1"
    os.makedirs(work, exist_ok=True)
    os.environ["ANIMEREC_ARTIFACT_DIR"] = work
    os.environ["ANIMEREC_KEEP_OUTPUTS"] = "1"
    os.chdir(work)
    from anime_recommendations_b200.components import neural_network, preprocess, similar_anime
    from oracle import similarity as osim
    out = dict(n_ratings=n)
    t0 = time.perf_counter()
    raw = synth_raw(n)
    raw.to_parquet("user_stats.parquet", index=False)
    n_anime = raw["anime_id"].nunique()
    pd.DataFrame({"MAL_ID": np.sort(raw["anime_id"].unique()), "Name": ["Anime %d" % i for i in range(n_anime)],
                  "English name": "e", "Japanese name": "j", "Score": 7.0, "Genres": "Action, Comedy", "Episodes": 12,
                  "Premiered": "Spring 2000", "Studios": "S", "Type": "TV", "Source": "Manga", "Rating": "PG",
                  "Members": 100}).to_csv("all_anime.csv", index=False)
    pd.DataFrame({"MAL_ID": np.sort(raw["anime_id"].unique()), "Name": ["Anime %d" % i for i in range(n_anime)],
                  "Genres": "Action, Comedy", "sypnopsis": "s"}).to_csv("synopses.csv", index=False)
    out["synth_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    pp = types.SimpleNamespace(raw_stats="user_stats.parquet:latest", project_name="p",
                               preprocessed_stats="preprocessed_stats.parquet", preprocessed_artifact_type="t",
                               preprocessed_artifact_description="d", num_reviews="400", drop_half_watched="False",
                               save_clean_locally="False", drop_unwatched="False", drop_plan="False")
    preprocess.go(pp)
    clean = pd.read_parquet("preprocessed_stats.parquet")
    out.update(preprocess_s=time.perf_counter() - t0, rows_after_preprocess=len(clean),
               users_after_filter=int(clean["user_id"].nunique()), min_ratings_per_user=int(clean["user_id"].value_counts().min()))

    cfg = dict(test_size="10000", TPU_INIT="False", embedding_size="128", kernel_initializer="he_normal",
               activation_function="sigmoid", model_loss="binary_crossentropy", optimizer="Adam", start_lr="0.00001",
               min_lr="0.00001", max_lr="0.00005", batch_size="10000", rampup_epochs="5", sustain_epochs="0",
               exp_decay="0.8", weights_artifact="wandb_main_weights.h5", save_weights_only="True",
               checkpoint_metric="val_loss", save_freq="epoch", mode="min", save_best_weights="True", verbose="0",
               epochs="1", save_model="True", model_name="./wandb_anime_nn.h5", input_data="preprocessed_stats.parquet:v2",
               project_name="p", model_artifact="wandb_anime_nn.h5", history_csv="wandb_anime_nn_history.csv",
               ID_emb_name="user_embedding", anime_emb_name="anime_embedding", merged_name="dot_product",
               main_df_type="t", model_type="h5", weights_type="h5", history_type="t", model_metrics='["mse"]',
               l2_reg_factor="0.0001", seed="1")
    argv = [x for k, v in cfg.items() for x in ("--" + k, v)]
    t0 = time.perf_counter()
    model, hist = neural_network.go(neural_network.parse(argv))
    out.update(train_s=time.perf_counter() - t0, epoch_device_s=model.timings["epoch_s"][-1],
               history={k: v[-1] for k, v in hist.history.items()}, n_users=model.n_users, n_anime=model.n_anime,
               train_samples_per_s=(len(clean) - 10000) / model.timings["epoch_s"][-1])

    t0 = time.perf_counter()
    sa = types.SimpleNamespace(main_df_type="t", anime_df_type="t", sypnopsis_df_type="t", model_type="h5",
                               model="wandb_anime_nn.h5:v12", project_name="p", main_df="preprocessed_stats.parquet:v2",
                               sypnopses_df="synopses.csv:v0", anime_df="all_anime.csv:v0", anime_query="Anime 123",
                               a_query_number="10", random_anime="False", anime_rec_genres="[None, None, None]",
                               an_spec_genres="False", types="['TV', 'Movie']", spec_types="True", a_rec_type="t",
                               save_sim_anime="True", ID_emb_name="user_embedding", anime_emb_name="anime_embedding")
    frame, fn = similar_anime.go(sa, model=model)
    out["similar_anime_s"] = time.perf_counter() - t0
    # parity spot check of the last stage against the reference's NumPy expressions on the saved weights
    W = model.get_layer("anime_embedding").get_weights()[0]
    a2i, ids = similar_anime.main_df_by_anime(sa)
    q = a2i[int(np.sort(raw["anime_id"].unique())[123])]
    oi, os_ = osim.similar_anime(W, q, 10)
    names = ["Anime %d" % int(np.searchsorted(np.sort(raw["anime_id"].unique()), ids[i])) for i in oi]
    out["similar_anime_matches_oracle"] = bool(frame["Name"].tolist() == names and
                                               np.allclose(frame["Similarity"].to_numpy(np.float32), os_, atol=3e-6))
    out["top3"] = frame["Name"].tolist()[:3]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
