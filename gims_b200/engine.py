"""Pair-parallel execution of the matcher on one GPU: several pairs in flight on separate CUDA streams.

Image pairs are independent (the reference loops over them one by one, eval_homography.py:161-177), so a
batch is issued round-robin over `n_streams` streams, each with its own workspace; across GPUs the batch is
simply split (`shard_pairs`) — there is no collective on the data path.
"""
import torch


def shard_pairs(n_pairs, world_size, rank):
    """Contiguous block of pair indices owned by `rank` (ceil split)."""
    per = (n_pairs + world_size - 1) // world_size
    lo = min(rank * per, n_pairs)
    return range(lo, min(lo + per, n_pairs))


class PairBatchRunner:
    """Runs batches of device-resident pairs through `GMatcher.run_pair` on a ring of streams."""

    def __init__(self, gmatcher, n_streams=4, pairs_per_launch=1):
        self.gm = gmatcher
        self.dev = gmatcher.bin_score.device
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams)]
        self.pairs_per_launch = max(1, int(pairs_per_launch))      # pairs stacked into one gims_forward_pairs call
        self._models = [gmatcher]

    def run(self, pairs, radius=25, percentile=7, min_size=8, keep=('matches0', 'mscores0', 'n_kept_dev')):
        """pairs: list of dicts with CUDA tensors keypoints{0,1} (N,2), descriptors{0,1} (D,N), scores{0,1} (N,),
        shape0/shape1.  Enqueues everything and returns per-pair dicts of the `keep` device tensors; the caller
        synchronises (e.g. `torch.cuda.current_stream().synchronize()` after `join()`)."""
        cur = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(cur)
        outs = []
        ppl = self.pairs_per_launch
        groups = [pairs[i:i + ppl] for i in range(0, len(pairs), ppl)]
        for i, grp in enumerate(groups):
            s = self.streams[i % len(self.streams)]
            if i < len(self.streams):
                s.wait_event(start)
            items = [(p['keypoints0'], p['descriptors0'], p['scores0'], p['keypoints1'], p['descriptors1'], p['scores1'],
                      p['shape0'], p['shape1']) for p in grp]
            with torch.cuda.stream(s):
                rs = self.gm.run_pairs(items, radius, percentile, min_size, stream=s, slot=i % len(self.streams))
            for r in rs:
                outs.append({k: r[k] for k in keep})
                self._last = r
        self.join()
        return outs

    def join(self):
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            e = torch.cuda.Event()
            e.record(s)
            cur.wait_event(e)
