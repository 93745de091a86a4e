"""CPU oracle for the anime_recommendations hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is shipped or measured as
the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

It restates, in NumPy float32 (with an optional float64 shadow), the arithmetic
the reference executes through TensorFlow-2.12/Keras and NumPy on the path named
in BASELINE.json ``north_star``:

* ``oracle.train``      -- neural_network/neural_network.py:66-125,184-217
* ``oracle.similarity`` -- similar_anime/similar_anime.py:136-171,399-468,
                           similar_users/similar_users.py:262-314,
                           model_recs/model_recs.py:132-192,373-456

PARITY UNPINNED (for the numeric results): the reference ships no tests and no
golden vectors for this path and TensorFlow is not
installable in this image, so the Keras-2.12 semantics encoded in
``oracle/train.py`` are assumptions, each listed in that module's docstring.
What IS pinned against the reference's own artefacts:

* ``lrfn``            vs the ``lr`` column of figure_file/anime_nn_history.csv
* data order          ``df.sample(frac=1, random_state=42)`` ==
                      ``np.random.RandomState(42).permutation(n)`` (golden made
                      with pandas in this container, tests/golden/)
* the manual backward vs PyTorch-CPU autograd of the same forward graph
* the similarity half is a line-for-line NumPy restatement of NumPy code, so it
  is the reference's own arithmetic (tests/golden/ fixtures were produced by
  executing the reference's expressions verbatim).
"""
from . import train, similarity  # noqa: F401
