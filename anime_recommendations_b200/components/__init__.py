"""Component entry points with the reference's argument names (SURVEY §8b "CLI surface to keep").

    python -m anime_recommendations_b200.components.preprocess      --raw_stats ... (preprocess/preprocess.py)
    python -m anime_recommendations_b200.components.neural_network  --test_size ... (neural_network/neural_network.py)
    python -m anime_recommendations_b200.components.similar_anime   --anime_query ... (similar_anime/similar_anime.py)
    python -m anime_recommendations_b200.components.similar_users   --sim_user_query ... (similar_users/similar_users.py)
    python -m anime_recommendations_b200.components.model_recs      --model_user_query ... (model_recs/model_recs.py)

W&B artifact references resolve to local files (components/_common.py:artifact_path); everything numeric
runs in libanimerec.so.
"""
