import torch


def _num_graphs(batch):
    return int(batch.max().item()) + 1 if batch.numel() else 0


def global_add_pool(x, batch, size=None):
    B = size or _num_graphs(batch)
    out = torch.zeros((B, x.shape[1]), dtype=x.dtype, device=x.device)
    return out.index_add_(0, batch, x)


def global_mean_pool(x, batch, size=None):
    B = size or _num_graphs(batch)
    cnt = torch.bincount(batch, minlength=B).clamp_min(1).to(x.dtype).unsqueeze(1)
    return global_add_pool(x, batch, B) / cnt


def global_max_pool(x, batch, size=None):
    B = size or _num_graphs(batch)
    out = torch.full((B, x.shape[1]), float("-inf"), dtype=x.dtype, device=x.device)
    idx = batch.unsqueeze(1).expand_as(x)
    return out.scatter_reduce(0, idx, x, reduce="amax", include_self=True)


class MessagePassing(torch.nn.Module):
    """aggr='add', flow='source_to_target' only (all the reference uses)."""

    def __init__(self, aggr="add"):
        super().__init__()
        assert aggr == "add"

    def propagate(self, edge_index, x, edge_attr):
        msg = self.message(x_j=x[edge_index[0]], edge_attr=edge_attr)
        out = torch.zeros((x.shape[0], msg.shape[1]), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, edge_index[1], msg)
        return self.update(out)

    def message(self, x_j, edge_attr):  # pragma: no cover - overridden
        return x_j

    def update(self, aggr_out):
        return aggr_out
