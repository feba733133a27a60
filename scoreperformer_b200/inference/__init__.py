"""Rendering loop and messengers (reference `scoreperformer.inference`)."""
from .generators import PerformanceData, ScorePerformerGenerator, render_performances
from .messengers import IntermediateData, SPMuple2IntermediateData, SPMuple2Messenger, SPMupleMessenger
from .token_tables import TokenTables
