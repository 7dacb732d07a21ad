"""Drop-in modules for the reference's model.py (jjery2243542/semi-supervised-ASR):
`pBLSTM`, `Encoder`, `AttLoc`, `Decoder`, `E2E`, `LM` with the same constructor arguments,
`forward` signatures, return values and `state_dict` key names/shapes (so reference `.ckpt`
files load unchanged), computing through the sm_100a kernels of liblas_b200.so.

Parameters live in stock `torch.nn` containers (`nn.LSTM`, `nn.LSTMCell`, `nn.Linear`, `nn.Conv2d`,
`nn.Embedding`) purely so that names, shapes and default initialisation match the reference; their
`forward` methods are never called. There is no CPU path: inputs are moved to the CUDA device and
the kernels raise if the extension is missing.
"""
import numpy as np
import torch

from . import functional as Fn
from .utils import _seq_mask, cc, pad_list, weight_init


class pBLSTM(torch.nn.Module):
    """model.py:58-98."""

    def __init__(self, input_dim, hidden_dim, n_layers, subsample, dropout_rate):
        super().__init__()
        layers, project_layers = [], []
        for i in range(n_layers):
            idim = input_dim if i == 0 else hidden_dim
            project_dim = hidden_dim * 4 if subsample[i] > 1 else hidden_dim * 2
            layers.append(torch.nn.LSTM(idim, hidden_dim, num_layers=1, bidirectional=True, batch_first=True))
            project_layers.append(torch.nn.Linear(project_dim, hidden_dim))
        self.layers = torch.nn.ModuleList(layers)
        self.project_layers = torch.nn.ModuleList(project_layers)
        self.dropout_layer = torch.nn.Dropout(p=dropout_rate)
        self.subsample = subsample
        self.dropout_rate = dropout_rate

    def _weights(self):
        w = []
        for layer, proj in zip(self.layers, self.project_layers):
            w += [layer.weight_ih_l0, layer.weight_hh_l0, layer.bias_ih_l0, layer.bias_hh_l0,
                  layer.weight_ih_l0_reverse, layer.weight_hh_l0_reverse, layer.bias_ih_l0_reverse,
                  layer.bias_hh_l0_reverse, proj.weight, proj.bias]
        return w

    def prep_jobs(self):
        """Weight-preparation jobs of this encoder for functional.prepare_ahead, in the order the step needs them."""
        w = self._weights()
        jobs = []
        for i in range(len(self.layers)):
            w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, proj_w, _ = w[10 * i:10 * i + 10]
            Dp = (w_ih.shape[1] + 7) // 8 * 8
            jobs += Fn.lstm_prep_jobs([w_ih, w_ih_r], [w_hh, w_hh_r], [b_ih, b_ih_r], [b_hh, b_hh_r], Dp)
            jobs.append(("cvt", [proj_w], (lambda pw: (lambda: Fn.cvt_bf16(pw)))(proj_w)))
        return jobs

    def forward(self, xpad, ilens):
        xpad = cc(xpad).float()
        host_lens = [int(l) for l in (ilens.tolist() if torch.is_tensor(ilens) else ilens)]
        T = max(host_lens)
        xpad = xpad[:, :T].contiguous()                      # pad_packed_sequence trims to the longest
        lens_dev = Fn.lens_tensor(host_lens, xpad.device)
        out = self.forward_dev(xpad, lens_dev)
        for sub in self.subsample:                            # model.py:92
            if sub > 1:
                host_lens = [(l + 1) // sub for l in host_lens]
        return out, host_lens

    def forward_dev(self, xpad, lens_dev):
        """Device-resident variant (no host work; CUDA-graph capturable): x f32 [B, T, D] with T the
        longest length, lens int32 [B] on the device. Returns enc_h only."""
        p = float(self.dropout_rate) if self.training else 0.0
        return Fn.EncoderFn.apply(xpad, lens_dev, tuple(int(s) for s in self.subsample), p, *self._weights())

    def forward_dev_split(self, xpad, lens_dev):
        """The same as forward_dev, as two autograd graphs: layer 0, then layers 1.. on a detached copy of its output.
        -> (enc_h, y0, y0_leaf): after `loss.backward()` the gradient sits in `y0_leaf.grad`, and
        `y0.backward(y0_leaf.grad)` runs layer 0's backward on its own -- the data-parallel trainer exchanges all other
        gradients while that last (and longest) BPTT kernel runs."""
        p = float(self.dropout_rate) if self.training else 0.0
        w = self._weights()
        subs = tuple(int(s) for s in self.subsample)
        y0 = Fn.EncoderFn.apply(xpad, lens_dev, subs[:1], p, *w[:10])
        if len(subs) == 1:
            return y0, None, None
        lens1 = lens_dev
        if subs[0] > 1:
            lens1 = torch.empty_like(lens_dev)
            Fn.call("las_pyramid_lens", Fn.ptr(lens_dev), lens_dev.numel(), subs[0], Fn.ptr(lens1))
        y0_leaf = y0.detach().requires_grad_(True)
        return Fn.EncoderFn.apply(y0_leaf, lens1, subs[1:], p, *w[10:]), y0, y0_leaf

    def out_lens_dev(self, lens_dev):
        """(len + 1) // sub per pyramid level, on the device (model.py:92)."""
        cur = lens_dev
        for sub in self.subsample:
            if sub > 1:
                nl = torch.empty_like(cur)
                Fn.call("las_pyramid_lens", Fn.ptr(cur), cur.numel(), int(sub), Fn.ptr(nl))
                cur = nl
        return cur


class Encoder(torch.nn.Module):
    """model.py:100-112."""

    def __init__(self, input_dim, hidden_dim, n_layers, subsample, dropout_rate, in_channel=1):
        super().__init__()
        self.enc2 = pBLSTM(input_dim=input_dim, hidden_dim=hidden_dim, n_layers=n_layers, subsample=subsample,
                           dropout_rate=dropout_rate)

    def forward(self, x, ilens):
        return self.enc2(x, ilens)


class AttLoc(torch.nn.Module):
    """model.py:114-173. Parameter container + per-utterance cache semantics; the attention math
    itself runs inside the fused decoder loop (functional.DecoderFn)."""

    def __init__(self, encoder_dim, decoder_dim, att_dim, conv_channels, conv_kernel_size, att_odim):
        super().__init__()
        self.mlp_enc = torch.nn.Linear(encoder_dim, att_dim)
        self.mlp_dec = torch.nn.Linear(decoder_dim, att_dim, bias=False)
        self.mlp_att = torch.nn.Linear(conv_channels, att_dim, bias=False)
        self.loc_conv = torch.nn.Conv2d(1, conv_channels, (1, 2 * conv_kernel_size + 1),
                                        padding=(0, conv_kernel_size), bias=False)
        self.gvec = torch.nn.Linear(att_dim, 1, bias=False)
        self.mlp_o = torch.nn.Linear(encoder_dim, att_odim)
        self.encoder_dim = encoder_dim
        self.decoder_dim = decoder_dim
        self.att_dim = att_dim
        self.att_odim = att_odim
        self.conv_channels = conv_channels
        self.conv_kernel_size = conv_kernel_size
        self.enc_length = None
        self.enc_h = None
        self.pre_compute_enc_h = None

    def reset(self):
        self.enc_length = None
        self.enc_h = None
        self.pre_compute_enc_h = None

    def _weights(self):
        return [self.mlp_enc.weight, self.mlp_enc.bias, self.mlp_dec.weight, self.mlp_att.weight,
                self.loc_conv.weight, self.gvec.weight, self.mlp_o.weight, self.mlp_o.bias]

    def _wdict(self):
        return dict(zip(("mlp_enc_w", "mlp_enc_b", "mlp_dec_w", "mlp_att_w", "conv_w", "gvec_w", "mlp_o_w", "mlp_o_b"),
                        self._weights()))

    def _precompute(self, enc_pad):
        """model.py:141-144: enc_h and mlp_enc(enc_h) are cached until reset()."""
        if self.pre_compute_enc_h is None:
            self.enc_h = enc_pad
            self.enc_length = enc_pad.size(1)
            self._enc_bf, P = Fn.attention_precompute(enc_pad, self._wdict())
            self.pre_compute_enc_h = P.view(enc_pad.size(0), enc_pad.size(1), -1)
        return self._enc_bf, self.pre_compute_enc_h.view(-1, self.att_dim)

    def forward(self, enc_pad, enc_len, dec_z, att_prev, scaling=2.0):
        """model.py:139-173, one attention read (inference-only; the training path runs the attention inside the
        fused decoder loop). -> (c [B, att_odim], w [B, Te])."""
        enc_pad = cc(enc_pad)
        B, Te, _ = enc_pad.shape
        enc_bf, P = self._precompute(enc_pad)
        W = dict(self._wdict(), emb_w=enc_pad.new_zeros(1, 1), w_hh=enc_pad.new_zeros(4 * self.decoder_dim, self.decoder_dim))
        return Fn.attention_step(W, enc_bf, P, B, Te, enc_len, dec_z, att_prev, self.conv_kernel_size, scaling)


class Decoder(torch.nn.Module):
    """model.py:256-367."""

    def __init__(self, output_dim, embedding_dim, hidden_dim, attention, att_odim, dropout_rate, bos, eos, pad,
                 ls_weight=0, labeldist=None):
        super().__init__()
        self.bos, self.eos, self.pad = bos, eos, pad
        self.embedding = torch.nn.Embedding(output_dim, embedding_dim, padding_idx=pad)
        self.LSTMCell = torch.nn.LSTMCell(embedding_dim + att_odim, hidden_dim)
        self.output_layer = torch.nn.Linear(hidden_dim + att_odim, output_dim)
        self.dropout_layer = torch.nn.Dropout(p=dropout_rate)
        self.attention = attention
        self.hidden_dim = hidden_dim
        self.att_odim = att_odim
        self.dropout_rate = dropout_rate
        self.ls_weight = ls_weight
        self.labeldist = labeldist
        self.vlabeldist = None
        if labeldist is not None:
            self.vlabeldist = cc(torch.from_numpy(np.array(labeldist, dtype=np.float32)))

    def zero_state(self, enc_pad, dim=None):
        return enc_pad.new_zeros(enc_pad.size(0), dim if dim else self.hidden_dim)

    def _weights(self):
        att = self.attention
        return [self.embedding.weight, self.LSTMCell.weight_ih, self.LSTMCell.weight_hh, self.LSTMCell.bias_ih,
                self.LSTMCell.bias_hh, self.output_layer.weight, self.output_layer.bias] + att._weights()

    def forward_step(self, emb, dec_z, dec_c, c, w, enc_pad, enc_len):
        """model.py:283-294, one decoder step with explicit state (inference-only; `Decoder.forward` runs whole
        sequences through the fused kernels). -> (logit, dec_z, dec_c, c, w)."""
        enc_pad = cc(enc_pad)
        B, Te, _ = enc_pad.shape
        enc_bf, P = self.attention._precompute(enc_pad)
        W = dict(zip(Fn.DEC_WEIGHTS, self._weights()))
        p = float(self.dropout_rate) if self.training else 0.0
        return Fn.decoder_step(W, enc_bf, P, B, Te, enc_len, cc(emb), dec_z, dec_c, c, w, self.attention.conv_kernel_size, p)

    def forward(self, enc_pad, enc_len, ys=None, tf_rate=1.0, max_dec_timesteps=500, sample=False, smooth=False,
                scaling=1.0, label_smoothing=True):
        dev = enc_pad.device
        B = enc_pad.size(0)
        self.attention.reset()
        enc_lens_dev = Fn.lens_tensor(enc_len, dev)
        teacher = ys is not None and len(ys) > 0
        tf_mask = None
        if teacher:
            ys_host = [y.detach().cpu().numpy().astype(np.int64) if torch.is_tensor(y) else np.asarray(y, np.int64)
                       for y in ys]
            L = max(len(y) for y in ys_host) + 1
            ys_in = np.full((B, L + 1), self.eos, dtype=np.int64)          # model.py:303-306
            ys_out = np.full((B, L), self.eos, dtype=np.int64)
            for b, y in enumerate(ys_host):
                ys_in[b, 0] = self.bos
                ys_in[b, 1:1 + len(y)] = y
                ys_out[b, :len(y)] = y
            ys_in[:, L] = self.pad                                          # guard column (never consumed)
            ys_in_dev = torch.from_numpy(ys_in).to(dev)
            ys_out_dev = torch.from_numpy(ys_out).to(dev)
            mode = 0
            if tf_rate < 1.0:
                # scheduled sampling (model.py:327-329): ONE host draw per step for the whole batch, from numpy's global
                # stream exactly as the reference draws it; step 0 always consumes <BOS>. The step loop then runs on the
                # free-running per-timestep kernels with the teacher's token substituted where the draw says so.
                draws = [np.random.random_sample() <= tf_rate for _ in range(L)]
                tf_mask = torch.tensor([1] + [int(d) for d in draws[1:]] + [1], dtype=torch.uint8, device=dev)
                mode = 1
        else:
            L = int(max_dec_timesteps)
            ys_in_dev, ys_out_dev = None, None
            mode = 2 if smooth else 1
        return self.forward_dev(enc_pad, enc_lens_dev, ys_in_dev, ys_out_dev, L, mode, scaling, label_smoothing,
                                tf_mask=tf_mask, sample=sample)

    def prep_jobs(self, mode=0):
        return Fn.decoder_prep_jobs(self._weights(), mode)

    def forward_dev(self, enc_pad, enc_lens_dev, ys_in_dev, ys_out_dev, L, mode, scaling=1.0, label_smoothing=True,
                    tf_mask=None, sample=False):
        """Device-resident variant (CUDA-graph capturable). ys_in_dev int64 [B, L+1] = [BOS, y, EOS.., PAD],
        ys_out_dev int64 [B, L] = [y, EOS..]; mode 0 teacher forcing, 1 greedy (with `tf_mask`: scheduled sampling
        between the teacher's tokens and the model's own predictions), 2 smooth free-run. `sample`: predictions are
        drawn from softmax(logits) instead of the argmax (model.py:349-351)."""
        dev = enc_pad.device
        p = float(self.dropout_rate) if self.training else 0.0
        in_kernel_sample = bool(sample) and mode == 1            # the only mode that feeds the prediction back
        weights = self._weights()
        # does a backward pass follow? (decides inference-only fast paths: one-launch greedy decoding, early stop)
        need_grad = torch.is_grad_enabled() and (enc_pad.requires_grad or any(w.requires_grad for w in weights))
        logits_alloc, ws_alloc, pred = Fn.DecoderFn.apply(
            enc_pad.float().contiguous(), enc_lens_dev, ys_in_dev, L, mode, float(scaling), 2.0,
            self.attention.conv_kernel_size, self.bos, p, tf_mask, in_kernel_sample, need_grad, *weights)
        ls = self.ls_weight if (label_smoothing and self.ls_weight > 0 and self.training) else 0.0
        dist = self.vlabeldist.to(dev) if ls > 0 else None
        targets, sampled = ys_out_dev, None
        if sample:
            # Categorical(logits).sample() (model.py:349-351): in-kernel where the prediction is fed back (mode 1), after
            # the loop where it is an output only; without targets the sampled tokens are what is scored (model.py:361)
            if mode == 1:
                sampled = pred
            else:
                V = logits_alloc.size(2)
                sampled = torch.multinomial(torch.softmax(logits_alloc[:, 1:].detach(), dim=-1).reshape(-1, V), 1).view(-1, L)
            if targets is None:
                targets = sampled
        ys_log_probs, _, prediction = Fn.CELabelSmoothFn.apply(logits_alloc, targets, dist, ls, 1, L)
        if sampled is not None:
            prediction = sampled
        return logits_alloc[:, 1:], ys_log_probs, prediction, ws_alloc[:, 1:]


class E2E(torch.nn.Module):
    """model.py:408-456."""

    def __init__(self, input_dim, enc_hidden_dim, enc_n_layers, subsample, dropout_rate, dec_hidden_dim, att_dim,
                 conv_channels, conv_kernel_size, att_odim, embedding_dim, output_dim, ls_weight, labeldist, pad=0,
                 bos=1, eos=2):
        super().__init__()
        self.encoder = Encoder(input_dim=input_dim, hidden_dim=enc_hidden_dim, n_layers=enc_n_layers,
                               subsample=subsample, dropout_rate=dropout_rate)
        self.attention = AttLoc(encoder_dim=enc_hidden_dim, decoder_dim=dec_hidden_dim, att_dim=att_dim,
                                conv_channels=conv_channels, conv_kernel_size=conv_kernel_size, att_odim=att_odim)
        self.decoder = Decoder(output_dim=output_dim, hidden_dim=dec_hidden_dim, embedding_dim=embedding_dim,
                               attention=self.attention, dropout_rate=dropout_rate, att_odim=att_odim,
                               ls_weight=ls_weight, labeldist=labeldist, bos=bos, eos=eos, pad=pad)

    def forward(self, data, ilens, ys=None, tf_rate=1.0, max_dec_timesteps=200, sample=False, smooth=False,
                scaling=1.0, label_smoothing=True):
        enc_h, enc_lens = self.encoder(data, ilens)
        return self.decoder(enc_h, enc_lens, ys, tf_rate=tf_rate, max_dec_timesteps=max_dec_timesteps,
                            sample=sample, smooth=smooth, scaling=scaling, label_smoothing=label_smoothing)

    def mask_and_cal_loss(self, log_probs, ys, mask=None):
        if mask is None:
            seq_len = [y.size(0) + 1 for y in ys]
            mask = _seq_mask(seq_len=seq_len, max_len=log_probs.size(1)).to(log_probs.device)
        else:
            seq_len = [y.size(0) for y in ys]
        return -torch.sum(log_probs * mask) / sum(seq_len)


class LM(torch.nn.Module):
    """The "judge" language model, model.py:459-573."""

    def __init__(self, output_dim, embedding_dim, hidden_dim, dropout_rate, n_layers, bos, eos, pad, ls_weight,
                 labeldist):
        super().__init__()
        self.bos, self.eos, self.pad = bos, eos, pad
        self.embedding = torch.nn.Embedding(output_dim, embedding_dim, padding_idx=pad)
        self.LSTM = torch.nn.LSTM(embedding_dim, hidden_dim, num_layers=n_layers, batch_first=True,
                                  dropout=dropout_rate if n_layers > 1 else 0)
        weight_init(self.LSTM)
        self.output_layer = torch.nn.Linear(hidden_dim, output_dim)
        self.dropout_layer = torch.nn.Dropout(p=dropout_rate)
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.dropout_rate = dropout_rate
        self.n_layers = n_layers
        self.ls_weight = ls_weight
        self.labeldist = labeldist
        self.vlabeldist = None
        if labeldist is not None:
            self.vlabeldist = cc(torch.from_numpy(np.array(labeldist, dtype=np.float32)))

    def zero_state(self, ref, dim=None):
        return ref.new_zeros(self.n_layers, ref.size(0), dim if dim else self.hidden_dim)

    def _weights(self):
        w = [self.embedding.weight]
        for l in range(self.n_layers):
            w += [getattr(self.LSTM, f"weight_ih_l{l}"), getattr(self.LSTM, f"weight_hh_l{l}"),
                  getattr(self.LSTM, f"bias_ih_l{l}"), getattr(self.LSTM, f"bias_hh_l{l}")]
        return w + [self.output_layer.weight, self.output_layer.bias]

    def targets(self, ys):
        """model.py:496-499: ys_in = [BOS, y], ys_out = [y, EOS], both EOS-padded to max(len) + 5; lens = len + 5.
        -> (ys_in, ys_out) int64 numpy [B, Lm], lens list[int]."""
        ys_host = [y.detach().cpu().numpy().astype(np.int64) if torch.is_tensor(y) else np.asarray(y, np.int64) for y in ys]
        B = len(ys_host)
        Lm = max(len(y) for y in ys_host) + 5
        ys_in = np.full((B, Lm), self.eos, dtype=np.int64)
        ys_out = np.full((B, Lm), self.eos, dtype=np.int64)
        for b, y in enumerate(ys_host):
            ys_in[b, 0] = self.bos
            ys_in[b, 1:1 + len(y)] = y
            ys_out[b, :len(y)] = y
        return ys_in, ys_out, [len(y) + 5 for y in ys_host]

    def forward(self, ys=None, discrete_input=True):
        dev = self.embedding.weight.device
        if discrete_input:
            ys_in, ys_out, lens = self.targets(ys)
            Lm = ys_in.shape[1]
            lens = Fn.lens_tensor(lens, dev)
            ys_in_dev = torch.from_numpy(ys_in).to(dev)
            ys_out_dev = torch.from_numpy(ys_out).to(dev)
        else:
            ys = ys.to(dev)
            bos_seq = ys.new_zeros(ys.size(0), 1) + self.bos                # model.py:503-505
            ys_in_dev = torch.cat([bos_seq, ys[:, :-1]], dim=1).contiguous()
            ys_out_dev = ys.contiguous()
            lens = None
            Lm = ys.size(1)
        return self.forward_dev(ys_in_dev, ys_out_dev, lens, Lm)

    def forward_dev(self, ys_in_dev, ys_out_dev, lens, Lm):
        """Device-resident variant (no host work; CUDA-graph capturable)."""
        dev = self.embedding.weight.device
        p = float(self.dropout_rate) if self.training else 0.0
        logits = Fn.LMFn.apply(ys_in_dev, lens, self.n_layers, self.pad, p, *self._weights())
        ls = self.ls_weight if (self.ls_weight > 0 and self.training) else 0.0
        dist = self.vlabeldist.to(dev) if ls > 0 else None
        ys_log_probs, ys_probs, predictions = Fn.CELabelSmoothFn.apply(logits, ys_out_dev, dist, ls, 0, Lm)
        return ys_log_probs, ys_probs, predictions

    def _step_ops(self):
        lw = [tuple(getattr(self.LSTM, f"{n}_l{l}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
              for l in range(self.n_layers)]
        return Fn.lm_step_operands(lw, self.output_layer.weight)

    def forward_step(self, emb, dec_z=None, dec_c=None, _ops=None):
        """model.py:535-542 ("only use in decode stage"): one LSTM timestep with explicit state + output layer.
        emb [B, 1, E]; dec_z / dec_c [n_layers, B, H] or None. -> (logit [B, V], dec_z, dec_c)."""
        emb = cc(emb)
        p = float(self.dropout_rate) if (self.training and self.n_layers > 1) else 0.0
        return Fn.lm_step(_ops or self._step_ops(), self.output_layer.bias, self.output_dim, emb.reshape(emb.size(0), -1),
                          dec_z, dec_c, p)

    @torch.no_grad()
    def decode(self, n_samples=5, sample=False, max_dec_timesteps=500):
        """model.py:544-563: free-running generation from <BOS>, greedy or sampled. -> predictions int64 [n, steps]."""
        dev = self.embedding.weight.device
        ops = self._step_ops()
        tok = torch.full((n_samples,), self.bos, device=dev, dtype=torch.int64)
        dec_z, dec_c, predictions = None, None, []
        for _ in range(max_dec_timesteps):
            emb = self.embedding.weight.index_select(0, tok).unsqueeze(1)
            logit, dec_z, dec_c = self.forward_step(emb, dec_z, dec_c, _ops=ops)
            if sample:
                tok = torch.multinomial(torch.softmax(logit, dim=-1), 1).squeeze(1)   # Categorical(logits).sample()
            else:
                tok = torch.argmax(logit, dim=-1)
            predictions.append(tok)
        return torch.stack(predictions, dim=1)

    def mask_and_cal_sum(self, log_probs, ys, mask=None):
        if mask is None:
            seq_len = [y.size(0) + 1 + 4 for y in ys]
            mask = _seq_mask(seq_len=seq_len, max_len=log_probs.size(1)).to(log_probs.device)
        else:
            seq_len = [y.size(0) for y in ys]
        return torch.sum(log_probs * mask) / sum(seq_len)
