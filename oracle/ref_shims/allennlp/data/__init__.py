"""Minimal `allennlp.data.Vocabulary` (non-padded namespace: @@UNKNOWN@@=0, @@BOUNDARY@@=1,
cf. reference var_updown/scripts/build_vocabulary.py:47,121-136)."""


class Vocabulary:
    def __init__(self, tokens=None):
        self._t2i = {}
        self._i2t = []
        for t in (tokens or ["@@UNKNOWN@@", "@@BOUNDARY@@"]):
            self.add_token_to_namespace(t)

    def add_token_to_namespace(self, token, namespace="tokens"):
        if token not in self._t2i:
            self._t2i[token] = len(self._i2t)
            self._i2t.append(token)
        return self._t2i[token]

    def get_vocab_size(self, namespace="tokens"):
        return len(self._i2t)

    def get_token_index(self, token, namespace="tokens"):
        return self._t2i.get(token, self._t2i["@@UNKNOWN@@"])

    def get_token_from_index(self, index, namespace="tokens"):
        return self._i2t[index]

    def get_token_to_index_vocabulary(self, namespace="tokens"):
        return dict(self._t2i)

    def get_index_to_token_vocabulary(self, namespace="tokens"):
        return dict(enumerate(self._i2t))
