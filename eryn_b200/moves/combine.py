"""CombineMove on the device (reference: moves/combine.py:11-135): the given moves in order, every one with its own
Metropolis step and tempering tail, composed on the host — the kernels are those of the sub-moves."""
import numpy as np

from .move import Move

__all__ = ["CombineMove"]


class CombineMove(Move):
    """Move that combines specific moves in order.

    Same constructor surface as eryn.moves.CombineMove: `moves` is a list of moves or (move, weight) tuples (weights
    are ignored, combine.py:16-18); `verbose` is accepted for compatibility (no progress bar on the device path)."""

    def __init__(self, moves, *args, verbose=False, **kwargs):
        self.moves = [m[0] if isinstance(m, tuple) else m for m in moves]
        if not self.moves:
            raise ValueError("CombineMove needs at least one move")
        self.verbose = verbose
        Move.__init__(self, *args, **kwargs)

    # ---- the counters live in the sub-moves (combine.py:32-48) ---------------------------------------------------
    @property
    def accepted(self):
        """accepted counts of each move (a list, as in the reference)"""
        return [move.accepted for move in self.moves]

    @accepted.setter
    def accepted(self, accepted):
        assert isinstance(accepted, np.ndarray)
        for move in self.moves:
            move.accepted = accepted.copy()

    @property
    def acceptance_fraction(self):
        """acceptance fraction averaged over all moves (combine.py:50-56)"""
        return np.mean([move.acceptance_fraction for move in self.moves], axis=0)

    @property
    def acceptance_fraction_separate(self):
        return [move.acceptance_fraction for move in self.moves]

    @property
    def temperature_control(self):
        return self._temperature_control

    @temperature_control.setter
    def temperature_control(self, temperature_control):  # combine.py:69-82
        for move in getattr(self, "moves", []):
            move.temperature_control = temperature_control
        self._temperature_control = temperature_control
        if temperature_control is not None:
            self.ntemps = temperature_control.ntemps

    @property
    def periodic(self):
        return self._periodic

    @periodic.setter
    def periodic(self, periodic):  # combine.py:89-97
        for move in getattr(self, "moves", []):
            move.periodic = periodic
        self._periodic = periodic

    @property
    def graphable(self):
        return all(getattr(m, "graphable", False) for m in self.moves)

    def _host_tick(self, n=1):
        for move in self.moves:
            move._host_tick(n)

    def bind(self, ctx):
        Move.bind(self, ctx)
        for move in self.moves:
            move.bind(ctx)

    def propose(self, model, state):
        """(state, accepted): `accepted` counts, per walker, the accepted proposals of all sub-moves (combine.py:99-135).
        A host State goes through every sub-move as a host State (as in the reference); a DeviceState stays resident."""
        accepted_out = None
        for move in self.moves:
            state, accepted = move.propose(model, state)
            if isinstance(accepted, np.ndarray):
                accepted_out = accepted.astype(np.int64) if accepted_out is None else accepted_out + accepted
            else:  # device mask (uint8, a scratch buffer the next sub-move overwrites)
                import torch
                accepted_out = accepted.to(torch.int32) if accepted_out is None else accepted_out + accepted
        return state, accepted_out
