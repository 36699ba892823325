class _Config:
    def update(self, *_a, **_k):
        pass


config = _Config()
