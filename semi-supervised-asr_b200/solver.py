"""`Solver` with the reference's entry points (solver.py of jjery2243542/semi-supervised-ASR) on the B200 path.

Same constructor (`Solver(config, load_model=False)`), same config.yaml keys, same public methods
(`sup_pretrain`, `judge_pretrain`, `ssl_train`, `validation`, `lm_validation`, `test`, `save_model` / `load_model`,
`save_judge` / `load_judge`), same checkpoint file names (`{model_dir}/{model_name}[-EEE].ckpt|.opt|.judge.ckpt|
.judge.opt`) and logger tags. The three train-step bodies (solver.py:288-301, 360-393, 460-495) are the
engines in engine.py; everything else here is host glue.

Data parallelism (new; the reference is single-process): when torch.distributed is initialised every rank builds
the same Solver, `batch_size` is the per-rank batch, batches are dealt by data.BatchLoader and gradients are
summed with one NCCL all-reduce per step (engine.py). Every rank decodes the (small) dev set itself -- all ranks
therefore agree on the best-CER bookkeeping without a broadcast -- and rank 0 alone logs and writes checkpoints.
"""
import os
import pickle

import numpy as np
import torch

from . import engine
from . import functional as Fn
from .data import BatchLoader, PickleDataset, collate, speech_collate, text_collate
from .model import E2E, LM
from .optim import FusedAdam
from .utils import (Logger, adjust_learning_rate, calculate_cer, cc, infinite_iter, load_resume_state, remove_pad_eos,
                    save_resume_state, to_gpu, to_sents)


def _dist():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


class Solver(object):
    def __init__(self, config, load_model=False):
        self.config = config
        self.rank, self.world = _dist()
        self.logger = Logger(config["logdir"]) if self.rank == 0 else None
        self.load_vocab()
        self.get_data_loaders()
        self.labeldist = self.get_label_dist(self.train_lab_dataset)             # solver.py:29
        self.unlab_labeldist = self.get_label_dist(self.train_unlab_y_dataset)   # solver.py:30
        self.proportion = self.calculate_length_proportion()                     # solver.py:33
        self.build_model(load_model=load_model)

    # ------------------------------------------------------------------ plumbing (solver.py:38-130)
    def log(self, kind, tag, value, step):
        if self.logger is not None:
            getattr(self.logger, kind)(tag, value, step)

    def say(self, *a, **k):
        if self.rank == 0:
            print(*a, **k)

    def save_model(self, model_path):
        if self.rank == 0:
            torch.save(self.model.state_dict(), f"{model_path}.ckpt")
            torch.save(self.gen_opt.state_dict(), f"{model_path}.opt")
            try:      # the sidecar is an addition to the reference's two files; it must never cost a checkpoint
                save_resume_state(f"{model_path}.resume", getattr(self, "progress", {}), Fn.current_dropout_seed(),
                                  loader_rng={n: getattr(self, n).rng_state() for n in self._TRAIN_LOADERS})
            except Exception as e:      # noqa: BLE001
                self.say(f"warning: resume sidecar not written ({e})")

    def save_judge(self, model_path):
        if self.rank == 0:
            torch.save(self.judge.state_dict(), f"{model_path}.judge.ckpt")
            torch.save(self.dis_opt.state_dict(), f"{model_path}.judge.opt")

    def load_vocab(self):
        with open(self.config["vocab_path"], "rb") as f:
            self.vocab = pickle.load(f)
        with open(self.config["non_lang_syms_path"], "rb") as f:
            self.non_lang_syms = pickle.load(f)

    def load_model(self, model_path, load_optimizer):
        self.model.load_state_dict(torch.load(f"{model_path}.ckpt"))
        if load_optimizer:
            self.gen_opt.load_state_dict(torch.load(f"{model_path}.opt"))
            # continuing a run (not just initialising from weights): counters and random streams from the sidecar,
            # when there is one (checkpoints written by the reference have none)
            state = load_resume_state(f"{model_path}.resume")
            if state is not None:
                self.progress = dict(state["progress"])
                if state.get("dropout_seed") is not None:
                    Fn.set_dropout_seed(state["dropout_seed"], next(self.model.parameters()).device)
                for n, rs in (state.get("loader_rng") or {}).items():
                    if hasattr(self, n):
                        getattr(self, n).set_rng_state(rs)

    def load_judge(self, model_path, load_optimizer):
        self.judge.load_state_dict(torch.load(f"{model_path}.judge.ckpt"))
        if load_optimizer:
            self.dis_opt.load_state_dict(torch.load(f"{model_path}.judge.opt"))

    def get_label_dist(self, dataset):
        """Unigram distribution of the targets incl. one <EOS> per utterance (solver.py:69-78)."""
        count = np.zeros(len(self.vocab))
        for _, y in dataset:
            np.add.at(count, np.asarray(y, dtype=np.int64), 1.0)
        count[self.vocab["<EOS>"]] += len(dataset)
        count[self.vocab["<PAD>"]] = 0
        count[self.vocab["<BOS>"]] = 0
        return count / np.sum(count)

    def calculate_length_proportion(self):
        """tokens per frame over the labeled set (solver.py:80-85)."""
        frames = sum(x.shape[0] for x, _ in self.train_lab_dataset)
        return sum(len(y) for _, y in self.train_lab_dataset) / frames

    _TRAIN_LOADERS = ("train_lab_loader", "train_unlab_x_loader", "train_unlab_y_loader")

    def get_data_loaders(self):
        c = self.config
        root = c["dataset_root_dir"]
        mk = lambda name, **kw: PickleDataset(os.path.join(root, f"{name}.pkl"), **kw)
        # not reference keys: `loader_prefetch: n` collates n batches ahead on a background thread into page-locked
        # buffers (the reference's DataLoader runs with num_workers=0); `bucket_batches: true` shuffles whole batches
        # of length-sorted utterances instead of utterances
        dp = dict(rank=self.rank, world=self.world, prefetch=int(c.get("loader_prefetch", 2)),
                  bucket=bool(c.get("bucket_batches", False)))
        self.train_lab_dataset = mk(c["labeled_set"], config=c, sort=True)
        self.train_lab_loader = BatchLoader(self.train_lab_dataset, c["batch_size"], c["shuffle"], False, collate, **dp)
        self.train_unlab_x_dataset = mk(c["unlabeled_speech_set"], config=c, sort=True)
        self.train_unlab_x_loader = BatchLoader(self.train_unlab_x_dataset, c["batch_size"], c["shuffle"], False,
                                                speech_collate, seed=1, **dp)
        self.train_unlab_y_dataset = mk(c["unlabeled_text_set"], config=c, sort=True)
        self.train_unlab_y_loader = BatchLoader(self.train_unlab_y_dataset, c["batch_size"], c["shuffle"], True,
                                                text_collate, seed=2, **dp)
        self.dev_dataset = mk(c["dev_set"], sort=True)                            # solver.py:118-123: no filters
        self.dev_loader = BatchLoader(self.dev_dataset, max(1, c["batch_size"] // 2), False, False, collate)

    def get_infinite_iter(self):
        self.lab_iter = infinite_iter(self.train_lab_loader)
        self.unlab_x_iter = infinite_iter(self.train_unlab_x_loader)
        self.unlab_y_iter = infinite_iter(self.train_unlab_y_loader)

    def build_model(self, load_model=False):
        c, v = self.config, self.vocab
        self.model = cc(E2E(input_dim=c["input_dim"], enc_hidden_dim=c["enc_hidden_dim"], enc_n_layers=c["enc_n_layers"],
                            subsample=c["subsample"], dropout_rate=c["dropout_rate"], dec_hidden_dim=c["dec_hidden_dim"],
                            att_dim=c["att_dim"], conv_channels=c["conv_channels"],
                            conv_kernel_size=c["conv_kernel_size"], att_odim=c["att_odim"], output_dim=len(v),
                            embedding_dim=c["embedding_dim"], ls_weight=c["ls_weight"], labeldist=self.labeldist,
                            pad=v["<PAD>"], bos=v["<BOS>"], eos=v["<EOS>"]))
        self.judge = cc(LM(output_dim=len(v), embedding_dim=c["dis_embedding_dim"], hidden_dim=c["dis_hidden_dim"],
                           dropout_rate=c["dis_dropout_rate"], n_layers=c["dis_layers"], bos=v["<BOS>"], eos=v["<EOS>"],
                           pad=v["<PAD>"], ls_weight=c["ls_weight"], labeldist=self.unlab_labeldist))
        if self.world > 1:                       # identical initial weights on every rank
            for p in list(self.model.parameters()) + list(self.judge.parameters()):
                torch.distributed.broadcast(p.data, 0)
        # solver.py:152-153, 171-173: Adam(amsgrad, wd) for the ASR, plain Adam for the judge
        self.gen_opt = FusedAdam(self.model.parameters(), lr=c["learning_rate"], weight_decay=c["weight_decay"],
                                 amsgrad=True)
        self.dis_opt = FusedAdam(self.judge.parameters(), lr=c["d_learning_rate"])
        if load_model:
            self.load_model(c["load_model_path"], c["load_optimizer"])
        self.sup_trainer = engine.SupervisedTrainer(self.model, self.gen_opt, max_grad_norm=c["max_grad_norm"],
                                                    use_graph=c.get("cuda_graph", True))
        self.ssl_trainer = engine.SSLTrainer(self.model, self.judge, self.gen_opt, max_grad_norm=c["max_grad_norm"],
                                             unsup_weight=c["unsup_weight"], proportion=self.proportion,
                                             smooth=c["smooth_embedding"], scaling=c["softmax_scaling"],
                                             # not a reference key: `guard_empty_mask: true` replaces solver.py:478's 0/0
                                             # (every free-run token == <EOS>) by 0; the default is the reference's NaN
                                             guard_empty_mask=bool(c.get("guard_empty_mask", False)))
        self.judge_trainer = engine.JudgeTrainer(self.judge, self.dis_opt, max_grad_norm=c["max_grad_norm"])
        if len(self.train_lab_dataset):
            # staging buffers sized once for the largest batch this run can draw: nothing is re-allocated per geometry
            Tmax = max(f.shape[0] for f, _ in self.train_lab_dataset)
            Lmax = max(len(y) for _, y in self.train_lab_dataset) + 1
            self.sup_trainer.reserve(c["batch_size"], Tmax, c["input_dim"], Lmax)

    # ------------------------------------------------------------------ evaluation (solver.py:176-286)
    def ind2sent(self, all_prediction, all_ys):
        hyp = to_sents(remove_pad_eos(all_prediction, eos=self.vocab["<EOS>"]), self.vocab, self.non_lang_syms)
        ref = to_sents(all_ys, self.vocab, self.non_lang_syms)
        return calculate_cer(hyp, ref), hyp, ref

    @torch.no_grad()
    def lm_validation(self):
        self.judge.eval()
        total = 0.0
        for data in self.dev_loader:
            _, _, ys = to_gpu(data)
            ys.sort(key=lambda y: len(y), reverse=True)
            log_probs, _, _ = self.judge(ys)
            total += float(-self.judge.mask_and_cal_sum(log_probs, ys))
        # solver.py:200-206: five sampled continuations from <BOS>, for the text summaries
        predictions = self.judge.decode(n_samples=5, sample=True, max_dec_timesteps=int(self.config.get("lm_decode_steps", 500)))
        sents = to_sents(remove_pad_eos(predictions.cpu().numpy().tolist(), eos=self.vocab["<EOS>"]), self.vocab,
                         self.non_lang_syms)
        self.judge.train()
        return total / max(1, len(self.dev_loader)), sents

    @torch.no_grad()
    def _decode_set(self, loader, with_loss):
        self.model.eval()
        preds, refs, total = [], [], 0.0
        # hypotheses are cut at their first <EOS> (ind2sent -> remove_pad_eos): decode in chunks and stop once
        # every utterance of the batch has produced one, instead of always running max_dec_timesteps steps
        Fn.GREEDY_EARLY_STOP.update(on=bool(self.config.get("greedy_early_stop", True)), eos=self.vocab["<EOS>"])
        try:
            return self._decode_loop(loader, with_loss, preds, refs, total)
        finally:
            Fn.GREEDY_EARLY_STOP["on"] = False

    def _decode_loop(self, loader, with_loss, preds, refs, total):
        for data in loader:
            xs, ilens, ys = to_gpu(data)
            if with_loss:
                _, log_probs, _, _ = self.model(xs, ilens, ys=ys)
                total += float(self.model.mask_and_cal_loss(log_probs, ys))
            _, _, prediction, _ = self.model(xs, ilens, ys=None, max_dec_timesteps=self.config["max_dec_timesteps"])
            preds += prediction.cpu().numpy().tolist()
            refs += [y.cpu().numpy().tolist() for y in ys]
        self.model.train()
        return preds, refs, total

    def validation(self):
        preds, refs, total = self._decode_set(self.dev_loader, with_loss=True)
        cer, hyp, ref = self.ind2sent(preds, refs)
        return total / max(1, len(self.dev_loader)), cer, hyp, ref

    def test(self, state_dict=None):
        if not state_dict:
            self.load_model(self.config["load_model_path"], self.config["load_optimizer"])
        else:
            self.model.load_state_dict(state_dict)
        test_set = self.config["test_set"]
        dataset = PickleDataset(os.path.join(self.config["dataset_root_dir"], f"{test_set}.pkl"), config=None, sort=False)
        preds, refs, _ = self._decode_set(BatchLoader(dataset, 1, False, False, collate), with_loss=False)
        cer, hyp, _ = self.ind2sent(preds, refs)
        if self.rank == 0:
            with open(f"{test_set}.txt", "w") as f:
                f.writelines(f"{p}\n" for p in hyp)
        self.say(f"{test_set}: {len(hyp)} utterances, CER={cer:.4f}")
        return cer

    # ------------------------------------------------------------------ training loops
    def _path(self, suffix=""):
        return os.path.join(self.config["model_dir"], self.config["model_name"]) + suffix

    def judge_train_one_iteration(self, unlab_ys):
        loss, avg_prob, _ = self.judge_trainer.step(unlab_ys)
        return {"loss": loss.item(), "avg_prob": avg_prob.item()}

    def judge_pretrain(self):
        c, tag = self.config, self.config["tag"]
        steps = len(self.train_unlab_y_loader)
        best = 100
        lr0 = c["d_learning_rate"]
        for epoch in range(c["judge_epochs"]):
            # MultiStepLR(milestones=[dis_change_learning_rate_epoch], gamma=lr_gamma), stepped at epoch start
            # (solver.py:308-315)
            adjust_learning_rate(self.dis_opt, lr0 * (c["lr_gamma"] if epoch + 1 >= c["dis_change_learning_rate_epoch"] else 1.0))
            total = 0.0
            for i, data in enumerate(self.train_unlab_y_loader):
                meta = self.judge_train_one_iteration([cc(y) for y in data])
                total += meta["loss"]
                for k, v in meta.items():
                    self.log("scalar_summary", f"{tag}/judge_pretrain/{k}", v, epoch * steps + i + 1)
            val_loss, _ = self.lm_validation()
            self.say(f"epoch: {epoch}, train_loss={total / max(1, steps):.3f}, valid_loss={val_loss:.3f}")
            self.log("scalar_summary", f"{tag}/judge_pretrain/val_loss", val_loss, epoch)
            self.log("scalar_summary", f"{tag}/judge_pretrain/avg_train_loss", total / max(1, steps), epoch)
            if val_loss < best:
                best = val_loss
                self.save_judge(self._path())
            self.save_judge(self._path(f"-{epoch:03d}"))

    def sup_train_one_epoch(self, epoch, tf_rate):
        c, tag = self.config, self.config["tag"]
        if tf_rate < 1.0:
            return self._sup_train_one_epoch_scheduled(epoch, tf_rate)
        steps = len(self.train_lab_loader)
        total = 0.0
        # solver.py:370-373: Gaussian input noise from `gaussian_epoch` on -- drawn on the device after the upload
        # (same distribution; the reference draws 8 M numpy normals per batch on the host)
        noisy = bool(c["add_gaussian"]) and epoch >= c["gaussian_epoch"]
        self.sup_trainer.input_noise_std = float(c["gaussian_std"]) if noisy else 0.0

        def batches():
            for xs, ilens, ys in self.train_lab_loader:
                yield xs, ilens, ys

        # pipelined: the H2D copy of batch i+1 overlaps step i; losses arrive as host scalars one step late, so
        # the GPU never waits for the host
        for i, (loss, _) in enumerate(self.sup_trainer.steps(batches())):
            total += float(loss)
            if self.logger is not None and (i + 1) % max(1, c.get("log_every", 50)) == 0:
                self.log("scalar_summary", f"{tag}/train_loss", loss.item(), epoch * steps + i + 1)
        return float(total) / max(1, steps)

    def _sup_train_one_epoch_scheduled(self, epoch, tf_rate):
        """solver.py:360-393 with tf_rate < 1 (scheduled sampling, model.py:327-329): the per-step choice between the
        teacher's token and the model's own prediction is a host draw, so these steps run eagerly through
        `E2E.forward(..., tf_rate=...)` (free-running per-timestep kernels) instead of the captured graph."""
        c, tag = self.config, self.config["tag"]
        steps = len(self.train_lab_loader)
        total = 0.0
        noisy = bool(c["add_gaussian"]) and epoch >= c["gaussian_epoch"]
        for i, (xs, ilens, ys) in enumerate(self.train_lab_loader):
            xs, ys = cc(xs), [cc(y) for y in ys]
            if noisy:
                xs = xs + torch.randn_like(xs) * float(c["gaussian_std"])
            self.model.train()
            _, log_probs, _, _ = self.model(xs, ilens, ys, tf_rate=tf_rate, sample=False)      # solver.py:375
            loss = -torch.mean(log_probs)
            self.gen_opt.zero_grad()
            with Fn.deferred_wgrad():
                loss.backward()
            engine._clip_and_step(self.gen_opt, list(self.model.parameters()), c["max_grad_norm"])
            total += float(loss)
            if self.logger is not None and (i + 1) % max(1, c.get("log_every", 50)) == 0:
                self.log("scalar_summary", f"{tag}/train_loss", float(loss), epoch * steps + i + 1)
        return total / max(1, steps)

    def sup_pretrain(self):
        c, tag = self.config, self.config["tag"]
        self.model.train()
        best_cer, best_model = 200, None
        first = 0
        prog = getattr(self, "progress", {})
        if c.get("resume") and prog.get("phase") == "sup_pretrain":      # `resume: true` is not a reference key
            first, best_cer = int(prog["epoch"]) + 1, float(prog.get("best_cer", best_cer))
        for epoch in range(first, c["epochs"]):
            if epoch <= c["tf_decay_epochs"]:                                     # solver.py:416-419
                tf_rate = c["init_tf_rate"] - (c["init_tf_rate"] - c["tf_rate_lowerbound"]) * (epoch / c["tf_decay_epochs"])
            else:
                tf_rate = c["tf_rate_lowerbound"]
            train_loss = self.sup_train_one_epoch(epoch, tf_rate)
            val_loss, cer, hyp, ref = self.validation()
            self.say(f"Epoch: {epoch}, tf_rate={tf_rate:.3f}, train_loss={train_loss:.4f}, valid_loss={val_loss:.4f}, "
                     f"CER={cer:.4f}")
            self.log("scalar_summary", f"{tag}/supervised/cer", cer, epoch)
            self.log("scalar_summary", f"{tag}/supervised/val_loss", val_loss, epoch)
            self.log("scalar_summary", f"{tag}/supervised/avg_train_loss", train_loss, epoch)
            for i, (p, g) in enumerate(zip(hyp[:5], ref[:5])):
                self.log("text_summary", f"{tag}/supervised/prediction-{i}", p, epoch)
                self.log("text_summary", f"{tag}/supervised/ground_truth-{i}", g, epoch)
            self.progress = {"phase": "sup_pretrain", "epoch": epoch, "best_cer": float(min(cer, best_cer))}
            if cer < best_cer:
                best_cer = cer
                self.save_model(self._path())
                best_model = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
            self.save_model(self._path(f"-{epoch:03d}"))
        return best_model, best_cer

    def gen_train_one_iteration(self, lab_xs, lab_ilens, lab_ys, unlab_xs, unlab_ilens):
        loss, sup, unsup, _ = self.ssl_trainer.step((lab_xs, lab_ilens, lab_ys), (unlab_xs, unlab_ilens))
        return {"unsup_loss": unsup.item(), "sup_loss": sup.item(), "loss": loss.item()}

    def ssl_train_one_iteration(self, iteration):
        lab, unlab = next(self.lab_iter), next(self.unlab_x_iter)
        lab_xs, lab_ilens, lab_ys = to_gpu(lab)
        meta = self.gen_train_one_iteration(lab_xs, lab_ilens, lab_ys, cc(unlab[0]), unlab[1])
        for k, v in meta.items():
            self.log("scalar_summary", f"{self.config['tag']}/ssl_generator/{k}", v, iteration + 1)
        return meta

    def ssl_train(self):
        c, tag = self.config, self.config["tag"]
        adjust_learning_rate(self.gen_opt, c["g_learning_rate"])                  # solver.py:519
        best_cer, best_model = 2, None
        if not hasattr(self, "lab_iter"):
            self.get_infinite_iter()
        total = c["ssl_iterations"]
        for step in range(total):
            self.ssl_train_one_iteration(step)
            if (step + 1) % c["summary_steps"] == 0 or step + 1 == total:
                val_loss, cer, hyp, ref = self.validation()
                self.say(f"Iter: [{step + 1}/{total}], valid_loss={val_loss:.4f}, CER={cer:.4f}")
                self.log("scalar_summary", f"{tag}/ssl/cer", cer, step + 1)
                self.log("scalar_summary", f"{tag}/ssl/val_loss", val_loss, step + 1)
                if cer < best_cer:
                    best_cer = cer
                    self.save_model(self._path())
                    self.save_judge(self._path())
                    best_model = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        return best_model, best_cer
