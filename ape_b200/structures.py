"""Minimal result containers with the detectron2 `Instances` / `Boxes` field surface the
reference's callers use (`instances.pred_boxes.tensor`, `.scores`, `.pred_classes`, `len()`,
`.to("cpu")`, `.image_size`).  When detectron2 is importable, `to_detectron2()` converts."""
import torch


class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor.reshape(-1, 4).to(torch.float32)

    def to(self, device):
        return Boxes(self.tensor.to(device))

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, item):
        return Boxes(self.tensor[item].view(-1, 4))


class Instances:
    def __init__(self, image_size, **fields):
        self._image_size = image_size
        self._fields = dict(fields)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self._fields[name] = val

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError(name)
        return self._fields[name]

    def has(self, name):
        return name in self._fields

    def get_fields(self):
        return self._fields

    def to(self, device):
        """Device -> host moves of large fields (pasted instance masks: 314 MB for 300 detections at 1024^2) go through pinned
        memory from PyTorch's caching host allocator (asynchronous copies at PCIe speed, one synchronisation for all fields)
        instead of a pageable copy per field."""
        if torch.device(device).type == "cpu":
            out, big = {}, False
            for k, v in self._fields.items():
                t = v.tensor if isinstance(v, Boxes) else v
                if torch.is_tensor(t) and t.is_cuda and t.numel() * t.element_size() >= (1 << 20):
                    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    host.copy_(t, non_blocking=True)
                    out[k], big = (Boxes(host) if isinstance(v, Boxes) else host), True
                else:
                    out[k] = v.to(device) if hasattr(v, "to") else v
            if big:
                torch.cuda.current_stream().synchronize()
            return Instances(self._image_size, **out)
        return Instances(self._image_size, **{k: (v.to(device) if hasattr(v, "to") else v) for k, v in self._fields.items()})

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def to_detectron2(self):
        from detectron2.structures import Boxes as D2Boxes
        from detectron2.structures import Instances as D2Instances

        out = D2Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, D2Boxes(v.tensor) if isinstance(v, Boxes) else v)
        return out
