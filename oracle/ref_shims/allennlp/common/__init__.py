class FromParams:
    pass


class Registrable(FromParams):
    default_implementation = None

    @classmethod
    def register(cls, name, **kwargs):
        def deco(sub):
            return sub
        return deco
