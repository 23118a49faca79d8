"""
Recursive update protocol shared by sources, boundaries and optical systems.

Same public behaviour as the reference's tfrt/update.py:3-78 (`update`, `forced_update`,
`update_handles`, `post_update_handles`, `recursively_update`, `frozen`); subclasses implement
`_update()` and `_generate_update_handles()`.
"""

from abc import ABC, abstractmethod


class RecursivelyUpdatable(ABC):
    def __init__(self, update_handles=None, recursively_update=True, frozen=False, **kwargs):
        self.recursively_update = recursively_update
        self.frozen = False   # the constructor's own update always runs (update.py:44-51)
        self.update_handles = (
            self._generate_update_handles() if update_handles is None else update_handles)
        self.post_update_handles = []
        self.update()
        self.frozen = frozen

    def update(self):
        """Recompute this object (and, recursively, what it consumes) unless frozen."""
        if self.frozen:
            return
        if self.recursively_update and self.update_handles:
            for handle in self.update_handles:
                handle()
        self._update()
        for handle in self.post_update_handles:
            handle()

    def forced_update(self):
        """Like update() but ignores `frozen`; does not force sub-objects (update.py:62-66)."""
        if self.recursively_update:
            for handle in self.update_handles:
                handle()
        self._update()

    @abstractmethod
    def _update(self):
        raise NotImplementedError

    @abstractmethod
    def _generate_update_handles(self):
        raise NotImplementedError
