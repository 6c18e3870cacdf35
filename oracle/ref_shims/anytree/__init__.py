class AnyNode:
    def __init__(self, parent=None, children=None, **kwargs):
        self.parent = parent
        self.children = children or []
        self.__dict__.update(kwargs)
