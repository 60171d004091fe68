"""stripedhyena.utils surface used by the reference (evo/models.py:8,142)."""


class dotdict(dict):
    """dict with attribute access; missing keys read as None (the reference builds it as
    ``dotdict(yaml_dict, Loader=...)``, evo/models.py:142, so extra kwargs become keys)."""

    def __getattr__(self, key):
        return self.get(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]
