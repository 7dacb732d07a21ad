"""End-to-end smoke of the Solver entry points (solver.py:303-358, 395-458, 516-565, 244-286) on synthetic
pickles in the reference's on-disk format: supervised pre-training, judge pre-training, SSL training, test()."""
import os
import pickle

import numpy as np
import pytest
import torch

from tests.util import pkg

pytestmark = pytest.mark.gpu


def _write_sets(root, rng, V):
    def mk(n, tmin, tmax):
        d = {}
        for i in range(n):
            T = int(rng.randint(tmin, tmax))
            d[f"utt{i:03d}"] = {"feature": rng.randn(T, 12).astype(np.float32),
                                "token_ids": [int(t) for t in rng.randint(3, V, size=max(2, T // 8))]}
        return d
    for name, n in (("lab", 12), ("unlab_x", 10), ("unlab_y", 12), ("dev", 4), ("test", 3)):
        with open(os.path.join(root, f"{name}.pkl"), "wb") as f:
            pickle.dump(mk(n, 24, 48), f)
    vocab = {"<PAD>": 0, "<BOS>": 1, "<EOS>": 2}
    for i in range(3, V):
        vocab[chr(ord("A") + i - 3)] = i
    with open(os.path.join(root, "vocab.pkl"), "wb") as f:
        pickle.dump(vocab, f)
    with open(os.path.join(root, "nls.pkl"), "wb") as f:
        pickle.dump(["<PAD>", "<BOS>", "<EOS>"], f)


def test_solver_entry_points(tmp_path):
    root = str(tmp_path)
    V = 10
    _write_sets(root, np.random.RandomState(0), V)
    cfg = dict(
        dataset_root_dir=root, vocab_path=os.path.join(root, "vocab.pkl"), non_lang_syms_path=os.path.join(root, "nls.pkl"),
        labeled_set="lab", unlabeled_speech_set="unlab_x", unlabeled_text_set="unlab_y", dev_set="dev", test_set="test",
        logdir=os.path.join(root, "log"), model_dir=root, model_name="m", load_model_path=os.path.join(root, "m"),
        load_optimizer=True, tag="t", max_feature_length=2300, min_feature_length=1, max_text_length=250,
        min_text_length=1, batch_size=4, shuffle=True, input_dim=12, enc_hidden_dim=16, enc_n_layers=2,
        subsample=[2, 2], dropout_rate=0.0, dec_hidden_dim=16, att_dim=16, conv_channels=3, conv_kernel_size=4,
        att_odim=16, embedding_dim=8, ls_weight=0.05, learning_rate=5e-3, weight_decay=1e-6, max_grad_norm=5,
        epochs=2, init_tf_rate=1.0, tf_rate_lowerbound=1.0, tf_decay_epochs=1, add_gaussian=False, gaussian_std=0.1,
        gaussian_epoch=0, max_dec_timesteps=8, dis_embedding_dim=8, dis_hidden_dim=16, dis_dropout_rate=0.0,
        dis_layers=2, d_learning_rate=2e-3, judge_epochs=2, dis_change_learning_rate_epoch=2, lr_gamma=0.2,
        g_learning_rate=1e-3, ssl_iterations=3, summary_steps=2, unsup_weight=0.001, smooth_embedding=True,
        softmax_scaling=3,
        # random synthetic text makes the tiny generator collapse onto <EOS>; the reference's unsupervised term is then
        # 0/0 (solver.py:478) and its NaN reaches every weight -- the opt-in guard (not a reference key) turns it into 0
        guard_empty_mask=True)
    S = pkg("solver")
    cwd = os.getcwd()
    os.chdir(root)
    try:
        s = S.Solver(cfg)
        assert abs(s.labeldist.sum() - 1) < 1e-12 and s.labeldist[0] == 0 and s.labeldist[1] == 0
        w0 = s.model.decoder.output_layer.weight.detach().clone()
        best, cer = s.sup_pretrain()
        assert best is not None and 0 <= cer
        assert not torch.equal(w0, s.model.decoder.output_layer.weight.detach())     # parameters moved
        s.judge_pretrain()
        s.ssl_train()
        for suffix in (".ckpt", ".opt", "-000.ckpt", "-001.opt", ".judge.ckpt", "-000.judge.opt"):
            assert os.path.exists(os.path.join(root, "m" + suffix)), suffix
        # checkpoints round-trip with the reference's key set (both attention.* aliases present)
        sd = torch.load(os.path.join(root, "m.ckpt"))
        assert any(k.startswith("attention.") for k in sd) and any(k.startswith("decoder.attention.") for k in sd)
        s2 = S.Solver(cfg, load_model=True)
        cer2 = s2.test()
        assert 0 <= cer2 and os.path.exists(os.path.join(root, "test.txt"))
        val_loss, cer3, hyp, ref = s2.validation()
        assert np.isfinite(val_loss) and len(hyp) == len(ref) == 4
        # resume sidecar: written next to the reference's files, restored with the optimiser,
        # and `resume: true` continues the supervised phase at the next epoch instead of starting over
        assert os.path.exists(os.path.join(root, "m.resume")) and os.path.exists(os.path.join(root, "m-001.resume"))
        assert s2.progress.get("phase") == "sup_pretrain" and s2.progress.get("epoch") == 1
        s3 = S.Solver(dict(cfg, resume=True, epochs=3), load_model=True)
        s3.sup_pretrain()
        assert s3.progress["epoch"] == 2 and os.path.exists(os.path.join(root, "m-002.ckpt"))
    finally:
        os.chdir(cwd)
