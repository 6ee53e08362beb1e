"""CPU restatement (numpy) of the hindsight-experience-replay store and 'future' relabelling - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench tools' checker legs may import this; the product path is the CUDA library.

What it restates: stable-baselines3 1.x `HerReplayBuffer` as the reference configures it
[REF benchmark/train.py:81-97: n_sampled_goal=4, goal_selection_strategy="future", max_episode_length=100,
online_sampling=True], i.e. (her_replay_buffer.py, SB3 1.x, restated from its published algorithm - stable-baselines3 is a
pip dependency of the reference [REF setup.py:17] that is neither vendored nor installed here):
  * `her_ratio = 1 - 1 / (n_sampled_goal + 1)`; the first `int(her_ratio * batch_size)` samples of a batch are relabelled;
  * an episode is drawn uniformly among the finished ones; relabelled samples of episodes longer than one transition draw
    the transition in [0, L-1) and the others in [0, L);
  * 'future': the new goal is `achieved_goal[episode, tf]` with tf uniform in [t+1, L);
  * `reward = env.compute_reward(next_achieved_goal, new_goal, info)` for the relabelled samples only; the others keep
    the stored reward; `done` is the stored `done * (1 - timeout)`.
PARITY UNPINNED against SB3 itself: there is no SB3 in this image and the reference holds no fixture of a sampled batch;
np.random could not be matched on a device in any case.  The parity contract is therefore this file <-> the CUDA path
(bit-exact: indices, gathered rows, relabelled goals, rewards), with the explicit index-drawing rule of include/xarm_abi.h
(Philox4x32-10, counter = (sample lo, sample hi, call, try), key = seed; range mapping by the high word of a 32x32-bit
product; rejection of unfinished episodes), and compute_reward pinned by the reward oracle (oracle/xarm_oracle.c).
Stated differences from SB3: per-env episode rings (SB3 1.x supports one env), one desired goal per episode.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
MAX_TRIES = 64


def philox4x32_10(c0, c1, c2, c3, seed):
    """Vectorised Philox4x32-10 (Random123): counters uint32 arrays, key = (seed lo, seed hi).  Returns four uint32 arrays."""
    c = [np.asarray(x, np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def mulhi(u, n):
    """floor(u * n / 2^32): the range mapping of the kernels (__umulhi)."""
    return ((np.asarray(u, np.uint64) * np.asarray(n, np.uint64)) >> np.uint64(32)).astype(np.int64)


class HerOracle:
    def __init__(self, num_envs, episodes_per_env, max_episode_length, obs_dim, goal_dim, action_dim, compute_reward,
                 n_sampled_goal=4, seed=0):
        N, K, T = num_envs, episodes_per_env, max_episode_length
        self.N, self.K, self.T, self.O, self.G, self.A = N, K, T, obs_dim, goal_dim, action_dim
        self.obs = np.zeros((K, T + 1, N, obs_dim), np.float32)
        self.ag = np.zeros((K, T + 1, N, goal_dim), np.float32)
        self.dg = np.zeros((K, N, goal_dim), np.float32)
        self.act = np.zeros((K, T, N, action_dim), np.float32)
        self.rew = np.zeros((K, T, N), np.float32)
        self.done = np.zeros((K, T, N), np.uint8)
        self.ep_len = np.zeros((K, N), np.int32)
        self.cur_k = np.zeros(N, np.int64)
        self.cur_t = np.zeros(N, np.int64)
        self.compute_reward = compute_reward      # f(next_achieved [n, G], desired [n, G]) -> [n] float32
        self.n_sampled_goal, self.seed, self.calls = n_sampled_goal, seed, 0
        self.episodes = self.transitions = 0

    def begin(self, obs, ag, dg, mask=None):
        i = np.arange(self.N) if mask is None else np.nonzero(np.asarray(mask))[0]
        k = self.cur_k[i]
        self.cur_t[i] = 0
        self.obs[k, 0, i], self.ag[k, 0, i], self.dg[k, i] = obs[i], ag[i], dg[i]

    def add(self, obs, ag, dg, terminal, action, reward, done, truncated=None):
        N, O, G = self.N, self.O, self.G
        i = np.arange(N)
        k, t = self.cur_k.copy(), self.cur_t.copy()
        d = np.asarray(done).astype(bool)
        nobs, nag = np.array(obs, np.float32), np.array(ag, np.float32)
        if terminal is not None:
            nobs[d], nag[d] = terminal[d, :O], terminal[d, O:O + G]
        self.obs[k, t + 1, i], self.ag[k, t + 1, i] = nobs, nag
        self.act[k, t, i], self.rew[k, t, i] = action, reward
        self.done[k, t, i] = d & ~(np.asarray(truncated).astype(bool) if truncated is not None else np.zeros(N, bool))
        close = d | (t + 1 == self.T)
        c = i[close]
        k2 = (k[c] + 1) % self.K
        self.obs[k2, 0, c], self.ag[k2, 0, c], self.dg[k2, c] = obs[c], ag[c], dg[c]
        self.ep_len[k[c], c] = t[c] + 1
        self.ep_len[k2, c] = 0
        self.cur_k[c], self.cur_t[c] = k2, 0
        self.cur_t[~close] += 1
        self.episodes += int(close.sum())
        self.transitions += N

    def sample_indices(self, batch):
        n_her = int((1.0 - 1.0 / (self.n_sampled_goal + 1)) * batch)
        b = np.arange(batch, dtype=np.uint64)
        idx = np.full((batch, 4), -1, np.int32)
        todo = np.ones(batch, bool)
        for tr in range(MAX_TRIES):
            if not todo.any():
                break
            bb = b[todo]
            c = philox4x32_10(bb & MASK, bb >> np.uint64(32), np.uint64(self.calls), np.uint64(tr), self.seed)
            env, k = mulhi(c[0], self.N), mulhi(c[1], self.K)
            L = self.ep_len[k, env].astype(np.int64)
            ok = L > 0
            her = (bb.astype(np.int64) < n_her) & (L > 1)
            Ls = np.maximum(L, 1)
            t = mulhi(c[2], np.where(her, Ls - 1, Ls))
            tf = np.where(her, t + 1 + mulhi(c[3], np.maximum(Ls - 1 - t, 0)), -1)
            rows = np.nonzero(todo)[0][ok]
            idx[rows] = np.stack([env, k, t, tf], 1)[ok]
            todo[rows] = False
        self.calls += 1
        return idx

    def sample(self, batch):
        idx = self.sample_indices(batch)
        ok = idx[:, 0] >= 0
        env, k, t, tf = (np.where(ok, idx[:, j], 0) for j in range(4))
        z = lambda a: np.where(ok.reshape(-1, *([1] * (a.ndim - 1))), a, 0).astype(a.dtype)
        her = ok & (idx[:, 3] >= 0)
        out = dict(index=idx)
        out["observation"], out["next_observation"] = z(self.obs[k, t, env]), z(self.obs[k, t + 1, env])
        out["achieved_goal"], out["next_achieved_goal"] = z(self.ag[k, t, env]), z(self.ag[k, t + 1, env])
        dg = self.dg[k, env].copy()
        dg[her] = self.ag[k[her], tf[her], env[her]]
        out["desired_goal"] = z(dg)
        out["action"] = z(self.act[k, t, env])
        rew = self.rew[k, t, env].copy()
        if her.any():
            rew[her] = self.compute_reward(out["next_achieved_goal"][her], out["desired_goal"][her])
        out["reward"], out["done"] = z(rew), z(self.done[k, t, env])
        return out
