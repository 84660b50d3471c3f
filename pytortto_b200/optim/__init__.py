from .optimizer import Optimizer
from .sgd import SGD
from .adam import Adam, AdamW
from . import lr_scheduler
