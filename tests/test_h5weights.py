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


def test_writer_round_trip_and_structure(tmp_path):
    """save_keras_weights -> a classic HDF5 file with the tree of the reference's own files; read back by the structural reader
    (which walks superblock / object headers / B-tree / symbol nodes / heaps like libhdf5) and by the layout-agnostic reader."""
    from cacto_b200.h5weights import load_keras_weights_by_tree, read_tree, save_keras_weights
    rng = np.random.default_rng(0)
    dims = [13, 64, 64, 128, 128, 1]
    w = []
    for i, o in zip(dims[:-1], dims[1:]):
        w += [rng.normal(size=(i, o)).astype(np.float32), rng.normal(size=o).astype(np.float32)]
    names = ['sinusodial_representation_dense', 'sinusodial_representation_dense_1', 'sinusodial_representation_dense_2',
             'sinusodial_representation_dense_3', 'dense_3']
    path = str(tmp_path / 'critic_5.h5')
    save_keras_weights(path, w, names)
    t = read_tree(path)
    assert t['attrs']['/'] == {'layer_names': names, 'backend': 'tensorflow', 'keras_version': '2.11.0'}
    assert t['attrs']['/dense_3']['weight_names'] == ['dense_3/kernel:0', 'dense_3/bias:0']
    assert t['datasets']['/dense_3/dense_3/kernel:0'].shape == (128, 1)
    for a, b in zip(load_keras_weights_by_tree(path), w):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(load_keras_weights(path, 13), w):
        np.testing.assert_array_equal(a, b)
    blob = open(path, 'rb').read()
    assert blob[:8] == b'\x89HDF\r\n\x1a\n' and int.from_bytes(blob[40:48], 'little') == len(blob)          # end-of-file address


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_structural_reader_walks_the_reference_files_and_rewrite_is_equivalent(tmp_path):
    """The reference's archived files parse with the same reader, in Keras' layer order; re-written by save_keras_weights they give
    the same arrays and the same (weight-bearing) layer names."""
    from cacto_b200.h5weights import load_keras_weights_by_tree, read_tree, save_keras_weights
    for net, fan_in in (('actor', 3), ('critic', 3), ('target_critic', 3)):
        src = os.path.join(REF, f'{net}_0.h5')
        t = read_tree(src)
        assert t['attrs']['/']['backend'] == 'tensorflow' and t['attrs']['/']['keras_version'] == '2.11.0'
        by_tree, by_shape = load_keras_weights_by_tree(src), load_keras_weights(src, fan_in)
        for a, b in zip(by_tree, by_shape):
            np.testing.assert_array_equal(a, b)
        names = [n for n in t['attrs']['/']['layer_names'] if t['attrs'].get('/' + n, {}).get('weight_names')]
        dst = str(tmp_path / f'{net}_0.h5')
        save_keras_weights(dst, by_tree, names)
        t2 = read_tree(dst)
        assert t2['attrs']['/']['layer_names'] == names
        assert sorted(t2['datasets']) == sorted(t['datasets'])
        for k in t['datasets']:
            np.testing.assert_array_equal(t2['datasets'][k], t['datasets'][k])
