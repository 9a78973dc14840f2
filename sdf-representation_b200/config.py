"""INI reader with the reference's keys (``configgen/config_reader.py:8-108``), instantiating the drop-in classes."""
import configparser

from . import losses, networks


class Configuration:
    def __init__(self, file_path="config.ini"):
        self.config = configparser.ConfigParser()
        if not self.config.read(file_path):
            raise FileNotFoundError(file_path)
        c = self.config
        self.geometry = c.get("Files", "geometry")
        self.directory = c.get("Files", "directory")
        self.name = c.get("Files", "name")
        name = c.get("Model", "model")
        if name != "ImplicitNet":
            raise ValueError(f"only the ImplicitNet hot path is implemented (got model = {name})")
        self.hidden_dim = c.getint("Model", "hidden_dim")
        self.num_hidden_layers = c.getint("Model", "num_hidden_layers")
        dims = [self.hidden_dim for _ in range(self.num_hidden_layers)]
        self.input_dim = c.getint("Model", "input_dim")
        val = c.getint("Model", "skip_connection")
        if val == 0:  # config_reader.py:26-29: no skip => ReLU + tanh output
            self.skip_connection = ()
            self.beta = 0
        else:
            self.skip_connection = (val,)
            self.beta = c.getfloat("Model", "beta")
        self.geometric_init = c.getboolean("Model", "geometric_init")
        self.model = networks.ImplicitNet(self.input_dim, dims, self.skip_connection, self.beta, self.geometric_init)
        self.loss = self.get_loss_function()
        self.lr = c.getfloat("Training", "lr")
        self.epochs = c.getint("Training", "epochs")
        self.minepochs = c.getint("Training", "min_epochs")
        self.batchsize = c.getint("Training", "batch_size")
        self.checkpointing = c.getint("Training", "checkpointing")
        self.contd = c.getboolean("Training", "continue")
        self.patience = c.getint("Training", "patience")
        self.two_dim = c.getboolean("Training", "two_dim") if c.has_option("Training", "two_dim") else False
        self.samplingonly = c.getboolean("Sampling", "samplingonly")
        self.continue_sampling = c.getboolean("Sampling", "continue_sampling")
        self.rescale = c.getboolean("Sampling", "rescale")
        self.distributed = c.getboolean("Sampling", "distributed")
        self.uniform_points = c.getint("Sampling", "uniform_points")
        self.surface = c.getint("Sampling", "surface")
        self.narrowband = c.getint("Sampling", "narrowband")
        self.narrowband_width = c.getfloat("Sampling", "narrowband_width")
        self.mismatchuse = c.getboolean("Sampling", "mismatchuse")
        self.train_test_split = c.getfloat("Sampling", "train_test_split")
        self.ppo = c.getboolean("Optional", "ppo")
        self.reconstruct = c.getboolean("Optional", "reconstruct")
        self.cubesize = c.getint("Optional", "cubesize")
        self.ppbatchsize = c.getint("Optional", "postprocessbatchsize")

    def get_loss_function(self):
        name = self.config.get("Loss", "loss_function")
        if not hasattr(losses, name):
            raise ValueError(f"Unsupported loss function: {name}")
        params = {k: float(self.config.get("Loss", k)) for k in self.config.options("Loss") if k != "loss_function"}
        return getattr(losses, name)(**params)
