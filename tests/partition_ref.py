"""TEST HELPER: torch restatement of the voting kernel of the GPU partitioner (lgcn_label_vote), pluggable into the
same orchestration (lgcn_b200/data/partition_gpu.py::partition(vote=...)) on any device."""
import torch


def torch_vote(ptr, nbr, labels, b, e, num_parts):
    """(want, best, own) for rows [b,e): most frequent out-neighbour label (ties: smallest), its count, the count of
    the row's own label."""
    ptr64, dev = ptr.long(), ptr.device
    deg = (ptr64[1:] - ptr64[:-1])[b:e]
    rows = torch.repeat_interleave(torch.arange(e - b, device=dev), deg)
    nb = nbr[ptr64[b]:ptr64[e]].long()
    cnt = torch.bincount(rows * num_parts + labels[nb], minlength=(e - b) * num_parts).view(e - b, num_parts)
    best = cnt.max(1).values
    want = cnt.argmax(1)                                   # first maximum = smallest label
    cur = labels[b:e]
    own = cnt.gather(1, cur[:, None]).squeeze(1)
    return torch.where(deg > 0, want, cur), torch.where(deg > 0, best, torch.zeros_like(best)), own
