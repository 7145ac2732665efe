class Mapping:
    def __init__(self, input_dim, output_dim, name="mapping"):
        self.input_dim = input_dim
        self.output_dim = output_dim

    def f(self, X):
        raise NotImplementedError
