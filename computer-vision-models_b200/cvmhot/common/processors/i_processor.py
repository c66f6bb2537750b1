"""Pre-processor plug-in interface (same contract as the reference's common/processors/i_processor.py:4-16)."""
from abc import ABC, abstractmethod


class IPreProcessor(ABC):
    """One sample at a time; processors are chained, each returning the 4-tuple it received."""

    @abstractmethod
    def process(self, raw_data, input_data, ground_truth, piped_params=None):
        return raw_data, input_data, ground_truth, piped_params
