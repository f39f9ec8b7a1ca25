from .observable import BaseObservable, LambdaObservable
from .operator import Observable
from .sigma_z import SigmaZ, AbsSigmaZ
