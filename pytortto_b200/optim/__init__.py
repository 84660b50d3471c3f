from .optimizer import Optimizer
from .sgd import SGD
from .adam import Adam
from . import lr_scheduler
