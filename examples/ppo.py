#!/usr/bin/env python
"""Device-resident PPO on the batched backend (the reference's counterpart is the agent glue under
pybulletgym/agents + pybulletgym/examples/tensorforce).

Everything stays on the GPU: VectorEnv.step() consumes and returns torch CUDA tensors, finished envs restart inside the
step kernel, so one PPO iteration is `horizon` kernel launches plus a handful of small matmuls.

    python examples/ppo.py --env HopperPyBulletEnv-v0 --envs 4096 --iters 150

Prints one line per iteration: env steps so far, mean return / length of the episodes that finished during the iteration
(device-side statistics of pbg_stats), steps per second including learning.
"""
import argparse
import math
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pybullet_gym_b200 import VectorEnv  # noqa: E402


class RunningNorm:
    def __init__(self, dim, device):
        self.mean = torch.zeros(dim, device=device)
        self.var = torch.ones(dim, device=device)
        self.count = 1e-4

    def update(self, x):
        b_mean, b_var, n = x.mean(0), x.var(0, unbiased=False), x.shape[0]
        delta = b_mean - self.mean
        tot = self.count + n
        self.mean = self.mean + delta * n / tot
        self.var = (self.var * self.count + b_var * n + delta * delta * self.count * n / tot) / tot
        self.count = tot

    def __call__(self, x):
        return ((x - self.mean) / torch.sqrt(self.var + 1e-8)).clamp(-10, 10)


class ActorCritic(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden=128):
        super().__init__()
        self.pi = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, act_dim))
        self.v = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, 1))
        self.log_std = nn.Parameter(torch.full((act_dim,), -0.5))
        for m in list(self.pi) + list(self.v):
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, math.sqrt(2)); nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.pi[-1].weight, 0.01)

    def dist(self, obs):
        return torch.distributions.Normal(self.pi(obs), self.log_std.exp())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="HopperPyBulletEnv-v0")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--iters", type=int, default=150)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--gamma", type=float, default=0.99)
    ap.add_argument("--lam", type=float, default=0.95)
    ap.add_argument("--clip", type=float, default=0.2)
    ap.add_argument("--ent", type=float, default=0.0, help="entropy bonus coefficient")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    torch.manual_seed(a.seed)
    dev = torch.device("cuda", 0)
    env = VectorEnv(a.env, a.envs, device=dev, seed=a.seed, auto_reset=True)
    E, T, D, A = a.envs, a.horizon, env.obs_dim, env.action_dim
    net = ActorCritic(D, A).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=a.lr, eps=1e-5)
    norm = RunningNorm(D, dev)
    obs = env.reset().clone()
    norm.update(obs)
    buf_o = torch.zeros(T, E, D, device=dev); buf_a = torch.zeros(T, E, A, device=dev)
    buf_lp = torch.zeros(T, E, device=dev); buf_r = torch.zeros(T, E, device=dev)
    buf_d = torch.zeros(T, E, device=dev); buf_v = torch.zeros(T + 1, E, device=dev)
    t0 = time.perf_counter()
    steps = 0
    for it in range(a.iters):
        env.stats(reset=True)
        with torch.no_grad():
            for t in range(T):
                no = norm(obs)
                d = net.dist(no)
                act = d.sample()
                buf_o[t], buf_a[t], buf_lp[t], buf_v[t] = no, act, d.log_prob(act).sum(-1), net.v(no).squeeze(-1)
                o2, r, done, _ = env.step(act)
                buf_r[t], buf_d[t] = r, done.float()
                obs = o2.clone()
            norm.update(obs)
            buf_v[T] = net.v(norm(obs)).squeeze(-1)
            adv = torch.zeros(T, E, device=dev)
            last = torch.zeros(E, device=dev)
            for t in reversed(range(T)):
                nd = 1.0 - buf_d[t]
                delta = buf_r[t] + a.gamma * buf_v[t + 1] * nd - buf_v[t]
                last = delta + a.gamma * a.lam * nd * last
                adv[t] = last
            ret = adv + buf_v[:T]
        steps += T * E
        fo, fa, flp, fadv, fret = buf_o.reshape(-1, D), buf_a.reshape(-1, A), buf_lp.reshape(-1), adv.reshape(-1), ret.reshape(-1)
        fadv = (fadv - fadv.mean()) / (fadv.std() + 1e-8)
        n = fo.shape[0]
        for ep in range(a.epochs):
            perm = torch.randperm(n, device=dev)
            for mb in perm.chunk(a.minibatches):
                d = net.dist(fo[mb])
                lp = d.log_prob(fa[mb]).sum(-1)
                ratio = (lp - flp[mb]).exp()
                pl = -torch.min(ratio * fadv[mb], ratio.clamp(1 - a.clip, 1 + a.clip) * fadv[mb]).mean()
                vl = 0.5 * (net.v(fo[mb]).squeeze(-1) - fret[mb]).pow(2).mean()
                loss = pl + 0.5 * vl - a.ent * d.entropy().sum(-1).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(net.parameters(), 0.5)
                opt.step()
        st = env.stats()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if st["episodes"]:
            print("iter %3d  steps %9d  episodes %6d  mean_return %8.1f  mean_len %6.1f  %.2e steps/s  (%.0f s)" % (
                it, steps, st["episodes"], st["return_sum"] / st["episodes"], st["length_sum"] / st["episodes"], steps / el, el), flush=True)


if __name__ == "__main__":
    main()
