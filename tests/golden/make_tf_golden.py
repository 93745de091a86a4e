"""Pin the Half-A oracle against the reference's real arithmetic -- run this WHERE TensorFlow 2.12 EXISTS
(`pip install tensorflow==2.12.0 numpy==1.23.5`; not possible in the build image: no wheel, no network).

It builds the model of neural_network.py:66-106 with the reference's own Keras calls, loads fixed initial weights,
runs a few `train_on_batch` steps at a fixed learning rate and dumps what TensorFlow computed:

    tests/golden/tf_train.npz   inputs (indices, labels, initial tables, head), per-step loss / mse, the tables,
                                the head, the BatchNorm moving statistics and the Adam slots after the last step

tests/test_oracle_train.py::test_oracle_matches_tensorflow_dump consumes the file when it is present (and is skipped,
saying "parity unpinned", when it is not), which turns the eight [K2.12] assumptions of oracle/train.py from
"read off the Keras source" into "checked against TensorFlow".  With --h5 PATH it also asks Keras to load a file
written by this repo's weights_io (the saved-weights layout check of SURVEY §8f-2).

    python tests/golden/make_tf_golden.py [--h5 wandb_anime_nn.h5]
"""
import argparse
import os
import sys

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))


def build(n_users, n_anime, dim, l2):
    import tensorflow as tf
    tfkl = tf.keras.layers
    reg = tf.keras.regularizers.L2(float(l2))
    user = tfkl.Input(name="user", shape=[1])
    ue = tfkl.Embedding(name="user_embedding", input_dim=n_users, output_dim=dim, embeddings_regularizer=reg)(user)
    anime = tfkl.Input(name="anime", shape=[1])
    ae = tfkl.Embedding(name="anime_embedding", input_dim=n_anime, output_dim=dim, embeddings_regularizer=reg)(anime)
    merged = tfkl.Dot(name="dot_product", normalize=True, axes=2)([ue, ae])
    merged = tfkl.Flatten()(merged)
    out = tfkl.Dense(1, kernel_initializer="he_normal")(merged)
    norm = tfkl.BatchNormalization()(out)
    act = tfkl.Activation("sigmoid")(norm)
    model = tf.keras.Model(inputs=[user, anime], outputs=act)
    model.compile(loss="binary_crossentropy", metrics=["mse"], optimizer="Adam")
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h5", default=None, help="a file written by anime_recommendations_b200.weights_io to load with Keras")
    args = ap.parse_args()
    import tensorflow as tf
    print("tensorflow", tf.__version__)
    n_users, n_anime, dim, B, steps, lr, l2 = 700, 90, 32, 256, 4, 1e-3, 1e-4
    rng = np.random.RandomState(123)
    U0 = rng.uniform(-0.05, 0.05, (n_users, dim)).astype(np.float32)
    A0 = rng.uniform(-0.05, 0.05, (n_anime, dim)).astype(np.float32)
    head0 = np.array([-1.3, 0.0, 1.0, 0.0], np.float32)                 # Dense kernel, bias, BN gamma, beta
    iu = rng.randint(0, n_users, (steps, B)).astype(np.int32)
    ia = rng.randint(0, n_anime, (steps, B)).astype(np.int32)
    ia[:, :40] = 3                                                       # a heavy row
    y = (rng.randint(0, 11, (steps, B)) / 10.0).astype(np.float32)
    model = build(n_users, n_anime, dim, l2)
    model.get_layer("user_embedding").set_weights([U0])
    model.get_layer("anime_embedding").set_weights([A0])
    model.get_layer("dense").set_weights([head0[0:1].reshape(1, 1), head0[1:2]])
    tf.keras.backend.set_value(model.optimizer.learning_rate, lr)
    loss, mse = [], []
    for s in range(steps):
        out = model.train_on_batch([iu[s], ia[s]], y[s], reset_metrics=True, return_dict=True)
        loss.append(out["loss"])
        mse.append(out["mse"])
    bn = model.get_layer("batch_normalization").get_weights()            # gamma, beta, moving_mean, moving_variance
    dense = model.get_layer("dense").get_weights()
    slots = {v.name: v.numpy() for v in model.optimizer.variables()}
    vu, va = rng.randint(0, n_users, 300), rng.randint(0, n_anime, 300)
    pred = model.predict([vu, va], verbose=0)
    np.savez_compressed(os.path.join(OUT, "tf_train.npz"), tf_version=np.array(tf.__version__), lr=lr, l2=l2,
                        U0=U0, A0=A0, head0=head0, iu=iu, ia=ia, y=y, loss=np.array(loss, np.float64),
                        mse=np.array(mse, np.float64), U=model.get_layer("user_embedding").get_weights()[0],
                        A=model.get_layer("anime_embedding").get_weights()[0],
                        head=np.array([dense[0][0, 0], dense[1][0], bn[0][0], bn[1][0]], np.float32),
                        moving=np.array([bn[2][0], bn[3][0]], np.float32), pred_u=vu, pred_a=va, pred=pred,
                        **{"slot__" + k.replace("/", "__"): v for k, v in slots.items()})
    print("wrote", os.path.join(OUT, "tf_train.npz"), "loss", loss)
    keras_file = os.path.join(OUT, "tf_model_keras.h5")
    model.save(keras_file)                                               # what the reference's model.save writes
    print("wrote", keras_file, "(input for tests/test_minih5.py::test_reads_a_keras_written_file)")
    if args.h5:
        m2 = tf.keras.models.load_model(args.h5)
        print("Keras loaded", args.h5, [w.shape for w in m2.get_weights()])
    return 0


if __name__ == "__main__":
    sys.exit(main())
