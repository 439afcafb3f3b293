# -*- coding: utf-8 -*-
"""Evaluator interface of lfd/evaluation/base_evaluator.py:6-13 (what Executor's EvaluationHook drives)."""

__all__ = ['Evaluator']


class Evaluator(object):

    def update(self, results):
        raise NotImplementedError

    def evaluate(self):
        raise NotImplementedError
