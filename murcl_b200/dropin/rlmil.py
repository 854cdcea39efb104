"""``Memory``, ``ActorCritic``, ``PPO`` and ``Full_layer`` of models/rlmil.py:7-239.

Same constructors, state-dict keys (``state_encoder.{0,2}``, ``gru.*_l0``, ``actor.0``, ``critic.0``;
``rnn.*_l0``, ``fc``) and call contracts.  The ``nn.GRU`` / ``nn.Linear`` members are parameter containers
(identical initialisation and checkpoint layout); the math runs on libmurcl_b200's dense-layer, GRU-cell
and actor-head kernels.  Hard-coded ``.cuda()`` calls of the reference become "the device of the input".
"""
import math

import torch
import torch.nn as nn

from .. import ops


class Memory:
    """Rollout buffer of one view (rlmil.py:7-22): parallel lists, one entry per patch-step."""

    FIELDS = ("actions", "states", "logprobs", "rewards", "is_terminals", "hidden")

    def __init__(self):
        for name in self.FIELDS:
            setattr(self, name, [])

    def clear_memory(self):
        for name in self.FIELDS:
            getattr(self, name).clear()


def _gru_params(gru: nn.GRU):
    return gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0


def _dtype(module):
    """Operand storage of the dense layers: fp32 (exact) or bf16 (tcgen05); module.precision or MURCL_PRECISION."""
    return ops.storage_dtype(getattr(module, "precision", None) or ops.default_precision())


class ActorCritic(nn.Module):
    def __init__(self, feature_dim, state_dim, hidden_state_dim=1024, policy_conv=False, action_std=0.1, action_size=2):
        super(ActorCritic, self).__init__()
        if policy_conv:
            # rlmil.py:30-37: 1x1 convolution over the feature map of the state, then one dense layer.  A 1x1 convolution is
            # a dense layer over the channels applied at every position, so it runs on the same GEMM kernels (_encode)
            self.state_encoder = nn.Sequential(
                nn.Conv2d(feature_dim, 32, kernel_size=1, stride=1, padding=0, bias=False), nn.ReLU(), nn.Flatten(),
                nn.Linear(int(state_dim * 32 / feature_dim), hidden_state_dim), nn.ReLU())
        else:
            self.state_encoder = nn.Sequential(
                nn.Linear(state_dim, 2048), nn.ReLU(),
                nn.Linear(2048, hidden_state_dim), nn.ReLU())
        self.gru = nn.GRU(hidden_state_dim, hidden_state_dim, batch_first=False)
        self.actor = nn.Sequential(nn.Linear(hidden_state_dim, action_size), nn.Sigmoid())
        self.critic = nn.Sequential(nn.Linear(hidden_state_dim, 1))
        self.action_std = float(action_std)        # the reference stores it as `action_var` but uses it as a std
        self.action_size = action_size
        self.hidden_state_dim = hidden_state_dim
        self.policy_conv = policy_conv
        self.feature_dim = feature_dim
        self.feature_ratio = int(math.sqrt(state_dim / feature_dim))

    def forward(self):
        raise NotImplementedError

    def _encode(self, state):
        dt = _dtype(self)
        if self.policy_conv:
            # state [N, F, r, r]: channels-last rows -> dense F -> 32 (+ReLU) -> back to the NCHW flattening order (rlmil.py:31-36)
            n, f, h, w = state.shape
            rows = state.permute(0, 2, 3, 1).reshape(n * h * w, f).contiguous()
            c = ops.linear(rows, self.state_encoder[0].weight.reshape(32, f), None, ops.ACT_RELU, dt)
            flat = c.reshape(n, h * w, 32).permute(0, 2, 1).reshape(n, 32 * h * w).contiguous()
            return ops.linear(flat, self.state_encoder[3].weight, self.state_encoder[3].bias, ops.ACT_RELU, dt)
        s = ops.linear(state, self.state_encoder[0].weight, self.state_encoder[0].bias, ops.ACT_RELU, dt)
        return ops.linear(s, self.state_encoder[2].weight, self.state_encoder[2].bias, ops.ACT_RELU, dt)

    def _state_rows(self, state_ini):
        """[N, ...] states as the encoder wants them: flattened vectors, or the [N, F, r, r] feature map (rlmil.py:71-74)."""
        return state_ini.float().contiguous() if self.policy_conv else state_ini.flatten(1).float().contiguous()

    def act(self, state_ini, memory, restart_batch=False, training=False, eps=None):
        """rlmil.py:66-97.  ``eps`` optionally supplies the standard-normal draw (parity tests); by default it
        is drawn with ``torch.randn`` on the state's device."""
        with torch.no_grad():
            if restart_batch:
                del memory.hidden[:]
                memory.hidden.append(torch.zeros(1, state_ini.size(0), self.hidden_state_dim, device=state_ini.device))
            enc = self._encode(self._state_rows(state_ini))
            h = ops.gru_step(enc, memory.hidden[-1][0].contiguous(), *_gru_params(self.gru), dtype=_dtype(self))
            memory.hidden.append(h.unsqueeze(0))
            logits = ops.linear(h, self.actor[0].weight, self.actor[0].bias)
            if eps is None:
                eps = torch.randn(logits.shape, device=logits.device, dtype=torch.float32)
            action, logprob, mean = ops.actor_head(logits, eps, self.action_std)
            if training:
                memory.states.append(state_ini)
                memory.actions.append(action)
                memory.logprobs.append(logprob)
            else:
                action = mean
        return action.detach()

    def act_views(self, states, memories, restart_batch=False, training=True, eps=None):
        """``act`` for several independent views in one batched pass (same math per view: rows do not interact).
        Used by the fused pre-training step; halves the number of small launches.  ``eps`` optionally supplies the
        standard-normal draws, one ``[B, K]`` tensor per view (parity tests)."""
        n, b = len(states), states[0].size(0)
        with torch.no_grad():
            if restart_batch:
                for m in memories:
                    del m.hidden[:]
                    m.hidden.append(torch.zeros(1, b, self.hidden_state_dim, device=states[0].device))
            enc = self._encode(torch.cat([self._state_rows(s) for s in states], 0).contiguous())
            h_prev = torch.cat([m.hidden[-1][0] for m in memories], 0).contiguous()
            h = ops.gru_step(enc, h_prev, *_gru_params(self.gru), dtype=_dtype(self))
            logits = ops.linear(h, self.actor[0].weight, self.actor[0].bias)
            if eps is None:
                eps = torch.randn(logits.shape, device=logits.device, dtype=torch.float32)
            else:
                eps = torch.cat([e.to(logits.device, torch.float32) for e in eps], 0)
            action, logprob, mean = ops.actor_head(logits, eps, self.action_std)
            outs = []
            for v, m in enumerate(memories):
                sl = slice(v * b, (v + 1) * b)
                m.hidden.append(h[sl].unsqueeze(0))
                if training:
                    m.states.append(states[v])
                    m.actions.append(action[sl])
                    m.logprobs.append(logprob[sl])
                    outs.append(action[sl].detach())
                else:
                    outs.append(mean[sl].detach())
        return outs

    def evaluate(self, state, action):
        """rlmil.py:99-127: log-prob, value and entropy of stored (state, action) sequences ``[T, B, ...]``."""
        seq_l, batch_size = state.size(0), state.size(1)
        if self.policy_conv:
            flat = state.reshape((seq_l * batch_size,) + tuple(state.shape[2:])).float().contiguous()     # rlmil.py:107
        else:
            flat = state.flatten(2).reshape(seq_l * batch_size, -1).float().contiguous()
        enc = self._encode(flat).reshape(seq_l, batch_size, -1)
        h = torch.zeros(batch_size, self.hidden_state_dim, device=state.device)
        outs = []
        for t in range(seq_l):
            h = ops.gru_step(enc[t].contiguous(), h, *_gru_params(self.gru), dtype=_dtype(self))
            outs.append(h)
        feat = torch.cat(outs, 0)
        mean = ops.linear(feat, self.actor[0].weight, self.actor[0].bias, ops.ACT_SIGMOID)
        value = ops.linear(feat, self.critic[0].weight, self.critic[0].bias)
        k, std = self.action_size, self.action_std
        a = action.reshape(seq_l * batch_size, -1)
        logprob = (-0.5 * (((a - mean) / std) ** 2).sum(1) - k * math.log(std) - 0.5 * k * math.log(2 * math.pi))
        entropy = torch.full_like(logprob, 0.5 * k * (1.0 + math.log(2 * math.pi)) + k * math.log(std))
        return logprob.view(seq_l, batch_size), value.view(seq_l, batch_size), entropy.view(seq_l, batch_size)


class PPO:
    def __init__(self, feature_dim, state_dim, hidden_state_dim, policy_conv,
                 action_std=0.1, lr=0.0003, betas=(0.9, 0.999), gamma=0.7, K_epochs=1, eps_clip=0.2, action_size=2):
        self.lr = lr
        self.betas = betas
        self.gamma = gamma
        self.eps_clip = eps_clip
        self.K_epochs = K_epochs

        self.policy = ActorCritic(feature_dim, state_dim, hidden_state_dim, policy_conv, action_std, action_size).cuda()
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=lr, betas=betas)
        self.policy_old = ActorCritic(feature_dim, state_dim, hidden_state_dim, policy_conv, action_std,
                                      action_size).cuda()
        self.policy_old.load_state_dict(self.policy.state_dict())
        self.MseLoss = nn.MSELoss()

    def select_action(self, state, memory, restart_batch=False, training=True):
        return self.policy_old.act(state, memory, restart_batch, training)

    def select_action_views(self, states, memories, restart_batch=False, training=True, eps=None):
        """``select_action`` for the two views of a patch-step (train_MuRCL.py:262-265) in one batched pass."""
        return self.policy_old.act_views(states, memories, restart_batch, training, eps=eps)

    def update(self, memory):
        """PPO-clip update (rlmil.py:152-184): discounted returns normalised over the whole rollout, K epochs of the
        clipped surrogate + 0.5 * value MSE - 0.01 * entropy, then policy_old <- policy."""
        returns, running = [], 0
        for reward in memory.rewards[::-1]:
            running = reward + self.gamma * running
            returns.append(running)
        returns = torch.cat(returns[::-1], 0).cuda()
        returns = (returns - returns.mean()) / (returns.std() + 1e-5)

        states, actions, old_logprobs = (torch.stack(seq, 0).cuda().detach()
                                         for seq in (memory.states, memory.actions, memory.logprobs))
        lo, hi = 1 - self.eps_clip, 1 + self.eps_clip
        for _ in range(self.K_epochs):
            logprobs, values, entropy = self.policy.evaluate(states, actions)
            ratio = (logprobs - old_logprobs).exp()
            advantage = returns - values.detach()
            surrogate = torch.min(ratio * advantage, ratio.clamp(lo, hi) * advantage)
            loss = 0.5 * self.MseLoss(values, returns) - surrogate - 0.01 * entropy
            self.optimizer.zero_grad()
            loss.mean().backward()
            self.optimizer.step()
        self.policy_old.load_state_dict(self.policy.state_dict())


class Full_layer(torch.nn.Module):
    def __init__(self, feature_num, hidden_state_dim=1024, fc_rnn=True, class_num=1000):
        super(Full_layer, self).__init__()
        self.class_num = class_num
        self.feature_num = feature_num
        self.hidden_state_dim = hidden_state_dim
        self.hidden = None
        self.fc_rnn = fc_rnn
        if fc_rnn:
            self.rnn = nn.GRU(feature_num, self.hidden_state_dim)
            self.fc = nn.Linear(self.hidden_state_dim, class_num)
        else:
            self.fc_2 = nn.Linear(self.feature_num * 2, class_num)
            self.fc_3 = nn.Linear(self.feature_num * 3, class_num)
            self.fc_4 = nn.Linear(self.feature_num * 4, class_num)
            self.fc_5 = nn.Linear(self.feature_num * 5, class_num)

    def forward_views(self, xs, restart=False):
        """``[self(x, restart) for x in xs]`` (train_MuRCL.py:243,272) with the view-independent dense layers batched:
        the input projection of all views is one GEMM, the recurrent part keeps the reference's single hidden chain
        (each call continues from the state the previous call left; ``restart`` resets it for every view), and the
        output layer is one GEMM over all views."""
        if not self.fc_rnn:
            return [self.forward(x, restart) for x in xs]
        n, b = len(xs), xs[0].size(0)
        dt = _dtype(self)
        w_ih, w_hh, b_ih, b_hh = _gru_params(self.rnn)
        gi_all = ops.linear(torch.cat([x.float() for x in xs], 0).contiguous(), w_ih, b_ih, ops.ACT_NONE, dt)
        hs = []
        h = None if restart or self.hidden is None else self.hidden[0]
        for v in range(n):
            h_prev = torch.zeros(b, self.hidden_state_dim, device=xs[0].device) if restart else h
            gh = ops.linear(h_prev, w_hh, b_hh, ops.ACT_NONE, dt)
            h = ops.gru_cell(gi_all[v * b:(v + 1) * b], gh, h_prev)
            hs.append(h)
        self.hidden = h.unsqueeze(0)
        out = ops.linear(torch.cat(hs, 0), self.fc.weight, self.fc.bias, ops.ACT_NONE, dt)
        return [out[v * b:(v + 1) * b] for v in range(n)]

    def forward(self, x, restart=False):
        if self.fc_rnn:
            # ``self.hidden`` is ONE state shared by every caller (both views of train_MuRCL.py:243,272 run
            # through it in turn) and is carried with its graph across the T patch-steps, as upstream.
            if restart:
                h_prev = torch.zeros(x.size(0), self.hidden_state_dim, device=x.device)
            else:
                h_prev = self.hidden[0]
            h = ops.gru_step(x.float().contiguous(), h_prev, *_gru_params(self.rnn), dtype=_dtype(self))
            self.hidden = h.unsqueeze(0)
            return ops.linear(h, self.fc.weight, self.fc.bias, ops.ACT_NONE, _dtype(self))
        else:
            if restart:
                self.hidden = x
            else:
                self.hidden = torch.cat([self.hidden, x], 1)
            width = self.hidden.size(1)
            if width == self.feature_num:
                return None
            for mult, layer in ((2, self.fc_2), (3, self.fc_3), (4, self.fc_4), (5, self.fc_5)):
                if width == self.feature_num * mult:
                    return ops.linear(self.hidden.float().contiguous(), layer.weight, layer.bias)
            raise RuntimeError(f"Full_layer(fc_rnn=False): unexpected accumulated width {width}")
