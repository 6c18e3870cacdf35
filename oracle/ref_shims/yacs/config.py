class CfgNode(dict):
    """Only so `updown.config` imports; the harness constructs the model directly."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
