from .i_processor import IPreProcessor
