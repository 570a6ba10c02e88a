"""cacto_b200.h5weights (Keras .h5 reader without h5py) against the reference's archived weight files.  Runs where
/root/reference exists (the build container); on the GPU box the committed fixtures cover the same arrays."""
import os

import numpy as np
import pytest

from cacto_b200.h5weights import chain_layers, load_keras_weights
from conftest import golden

REF = '/root/reference/Results Single Integrator/Results set test/NNs/N_try_0'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_reads_reference_h5_files():
    g = golden('h5_si_try0.npz')
    for net, n in (('actor', 6), ('critic', 10), ('target_critic', 10)):
        w = load_keras_weights(os.path.join(REF, f'{net}_0.h5'), 3)
        assert len(w) == n
        for i, a in enumerate(w):
            np.testing.assert_array_equal(a, g[f'{net}_{i}'])
    # shapes and initialiser ranges of NeuralNetwork.py:51-63 / :95-108
    actor = load_keras_weights(os.path.join(REF, 'actor_0.h5'), 3)
    assert [a.shape for a in actor] == [(3, 256), (256,), (256, 256), (256,), (256, 2), (2,)]
    critic = load_keras_weights(os.path.join(REF, 'critic_0.h5'), 3)
    assert [a.shape for a in critic] == [(3, 64), (64,), (64, 64), (64,), (64, 128), (128,), (128, 128), (128,), (128, 1), (1,)]
    assert np.abs(critic[0]).max() <= np.sqrt(6 / 3) and np.abs(critic[0]).max() > 1.3
    assert np.all(actor[1] == 0) and np.all(critic[9] == 0)


def test_chain_layers_disambiguates_equal_sizes():
    rng = np.random.default_rng(0)
    dims = [5, 64, 64, 128, 128, 1]
    layers = [(rng.normal(size=i * o).astype(np.float32), rng.normal(size=o).astype(np.float32)) for i, o in zip(dims[:-1], dims[1:])]
    shuffled = [layers[4], layers[0], layers[1], layers[2], layers[3]]          # Keras order: dense_N first
    out = chain_layers(shuffled, 5)
    for l, (k, b) in enumerate(layers):
        np.testing.assert_array_equal(out[2 * l].reshape(-1), k)
        np.testing.assert_array_equal(out[2 * l + 1], b)
    with pytest.raises(ValueError):
        chain_layers(shuffled, 7)
