import torch


class Data:
    """Attribute bag.  `.to(device)` moves tensor attributes only (so a pandas `pos_df` stays put),
    which is what `gnn_datasets.py:19-20` relies on."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    def to(self, device, *args, **kwargs):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, *args, **kwargs))
        return self

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={type(v).__name__}")
        return "Data(" + ", ".join(parts) + ")"
