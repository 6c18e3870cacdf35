import torch


class Vectors:
    def __init__(self, name=None, cache=None, **kwargs):
        self.stoi = {}
        self.vectors = torch.zeros(0, 300)


class GloVe(Vectors):
    def __init__(self, name="840B", dim=300, **kwargs):
        super().__init__(name=name, **kwargs)
