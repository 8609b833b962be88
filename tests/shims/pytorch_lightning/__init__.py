"""Minimal stand-in for pytorch_lightning: LightningModule = nn.Module with the hooks model_pl.py defines, Trainer.test = move to the
device, eval, inference_mode, loop test_step over test_dataloader() (TEST INFRASTRUCTURE)."""
import torch


class LightningModule(torch.nn.Module):
    global_rank = 0

    def test_dataloader(self):
        return None


def _to_device(x, dev):
    if torch.is_tensor(x):
        return x.to(dev)
    if isinstance(x, dict):
        return {k: _to_device(v, dev) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_device(v, dev) for v in x)
    return x


class Trainer:
    def __init__(self, devices=None, accelerator="auto", **kw):
        self.devices, self.accelerator = devices, accelerator

    def test(self, model, dataloaders=None):
        use_cuda = self.accelerator in ("cuda", "gpu", "auto") and torch.cuda.is_available()
        dev = torch.device("cuda", (self.devices or [0])[0]) if use_cuda else torch.device("cpu")
        model.to(dev)
        model.eval()
        loader = dataloaders if dataloaders is not None else model.test_dataloader()
        with torch.inference_mode():
            for i, batch in enumerate(loader):
                model.test_step(_to_device(batch, dev), i)
            if hasattr(model, "on_test_epoch_end"):
                model.on_test_epoch_end()
        return []
