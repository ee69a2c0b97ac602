"""Mini-batch assembly for training, on the device: the data path of GenericNNetWrapper.train.

What it replaces (reference, GenericNNetWrapper.py):
  * compute_surprise_weights :333-341    w = s / sum(s) + 1 / N, renormalised
  * the sampling of a batch      :68     np.random.choice(N, batch_size, replace=False, p=weights or None)
  * pick_examples + target build :69-80  boards -> float32, valids -> bool, pi / v -> float32 and the score-difference
    targets: a one-hot float32[B, 2 * max_diff + 1, n] with a 1 at clip(scdiff + max_diff, 0, 2 * max_diff) per player

The examples stay where batched self-play left them (device tensors, examples.FIELDS); nothing is pickled, decompressed
or copied through the host per batch. Optimiser, losses and the network's training mode stay the reference's
(GenericNNetWrapper.train feeds on these tensors unchanged). Pure tensor plumbing (torch), no custom kernel: per batch it is
one multinomial, six gathers and one scatter.

Note on `surprise`: the reference stores getActionProb's second return value (the q vector, length n) in that slot
(Coach.py:75,80), which its own compute_surprise_weights cannot digest (np.random.choice needs 1-d p); weights are therefore
taken from a per-example scalar the caller chooses (default: |q[0]|).
"""
import torch


def surprise_weights(surprise):
    """compute_surprise_weights (:333-341) for a per-example scalar float[N] -> float64[N] sampling probabilities"""
    s = surprise.to(torch.float64).reshape(-1)
    w = s / s.sum() + 1.0 / s.numel()
    return w / w.sum()


def scdiff_targets(scdiff, max_diff=15):
    """:76-80 -> float32[B, 2 * max_diff + 1, n] one-hot over the clipped score difference of every player"""
    B, n = scdiff.shape
    idx = (scdiff.to(torch.int64) + max_diff).clamp_(0, 2 * max_diff)
    out = torch.zeros((B, 2 * max_diff + 1, n), dtype=torch.float32, device=scdiff.device)
    out.scatter_(1, idx.view(B, 1, n), 1.0)
    return out


class TrainBatcher:
    """examples: dict of tensors (examples.FIELDS) on one device. `batch()` returns what one iteration of the reference's
    inner training loop builds (:68-84): boards float32[B,R,7], valid_actions bool[B,406], target_pis float32[B,406],
    target_vs float32[B,n], target_scdiffs float32[B,31,n], and the sample ids."""

    def __init__(self, examples, batch_size, max_diff=15, surprise_weight=False, surprise_scalar=None, seed=None):
        self.ex = examples
        self.N = int(examples["board"].shape[0])
        self.batch_size = int(batch_size)
        self.max_diff = int(max_diff)
        if self.batch_size > self.N:
            raise ValueError(f"batch_size {self.batch_size} > {self.N} examples (the reference samples without replacement)")
        dev = examples["board"].device
        self.gen = torch.Generator(device=dev)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        self.weights = None
        if surprise_weight:
            s = surprise_scalar if surprise_scalar is not None else examples["surprise"][:, 0].abs()
            self.weights = surprise_weights(s)

    @property
    def batches_per_epoch(self):
        return self.N // self.batch_size      # :54

    def sample_ids(self):
        if self.weights is None:              # uniform, without replacement
            return torch.randperm(self.N, device=self.ex["board"].device, generator=self.gen)[: self.batch_size]
        return torch.multinomial(self.weights, self.batch_size, replacement=False, generator=self.gen)

    def batch(self, ids=None):
        ids = self.sample_ids() if ids is None else ids
        ex = self.ex
        return dict(boards=ex["board"][ids].to(torch.float32), valid_actions=ex["valids"][ids].to(torch.bool),
                    target_pis=ex["pi"][ids].to(torch.float32), target_vs=ex["winner"][ids].to(torch.float32),
                    target_scdiffs=scdiff_targets(ex["scdiff"][ids], self.max_diff), ids=ids)
