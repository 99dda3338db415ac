import torch


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def to_dense_batch(x, batch, max_num_nodes=None):
    B = int(batch.max().item()) + 1
    counts = torch.bincount(batch, minlength=B)
    N = int(max_num_nodes or counts.max().item())
    start = torch.cumsum(counts, 0) - counts
    pos = torch.arange(batch.numel(), device=batch.device) - start[batch]
    out = x.new_zeros((B, N) + tuple(x.shape[1:]))
    mask = torch.zeros((B, N), dtype=torch.bool, device=x.device)
    out[batch, pos] = x
    mask[batch, pos] = True
    return out, mask


def to_dense_adj(edge_index, batch, edge_attr, max_num_nodes=None):
    B = int(batch.max().item()) + 1
    counts = torch.bincount(batch, minlength=B)
    N = int(max_num_nodes or counts.max().item())
    start = torch.cumsum(counts, 0) - counts
    b = batch[edge_index[0]]
    i = edge_index[0] - start[b]
    j = edge_index[1] - start[b]
    adj = edge_attr.new_zeros((B, N, N) + tuple(edge_attr.shape[1:]))
    adj.index_put_((b, i, j), edge_attr, accumulate=True)
    return adj
