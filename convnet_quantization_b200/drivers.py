"""Protocol restatement of the reference's two drivers, for use where /root/reference is not available (GPU box).

``InferenceBenchmark`` follows ``utils/inference_benchmark.py:10-157`` (warm-up 10 forwards; 100 timed forwards;
img/s = batch*iters / sum of wall times; compare batch 32) and ``ModelEvaluator`` follows
``utils/model_evaluator.py:11-55`` (``model.cpu()``, CPU images, top-1/top-5 via ``topk``).  Same class and method
names, argument meaning and return values, so code written against the reference's drivers runs against these.
The one deliberate difference: when ``device`` is CUDA, each timed forward is bracketed by a device synchronize —
the reference reads ``time.time()`` without one (``:95-98``), which on a GPU times the launch, not the work.
The reference's own driver files also drive the models of this package unchanged (same duck-typed surface).
"""
from __future__ import annotations

import time

import numpy as np
import torch


def _first_batch(loader):
    images, _ = next(iter(loader))
    return images


def _sync(device):
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()


class InferenceBenchmark:
    def __init__(self, test_loader, device="cpu"):
        self.test_loader = test_loader
        self.device = device

    def _prepare(self, model):
        model.eval()
        model.to(self.device)  # result intentionally discarded, as in the reference (:20)

    def _time_forwards(self, model, data, num_iterations):
        times = []
        with torch.no_grad():
            for _ in range(num_iterations):
                _sync(self.device)
                t0 = time.time()
                model(data)
                _sync(self.device)
                times.append(time.time() - t0)
        return times

    def warm_up(self, model, num_iterations=10):
        self._prepare(model)
        data = _first_batch(self.test_loader).to(self.device)
        with torch.no_grad():
            for _ in range(num_iterations):
                model(data)
        _sync(self.device)

    def measure_inference_time(self, model, batch_size=1, num_iterations=100, verbose=True):
        self._prepare(model)
        batch = _first_batch(self.test_loader)
        single = np.array(self._time_forwards(model, batch[:1].to(self.device), num_iterations)) * 1e3
        many = np.array(self._time_forwards(model, batch[:batch_size].to(self.device), num_iterations)) * 1e3
        result = {"single": (float(single.mean()), float(single.std())),
                  "batch": (float(many.mean()), float(many.std())),
                  "per_image": float(many.mean()) / batch_size}
        if verbose:
            print(f"single image: {result['single'][0]:.3f} +- {result['single'][1]:.3f} ms; "
                  f"batch {batch_size}: {result['batch'][0]:.3f} ms ({result['per_image']:.4f} ms/image)")
        return result

    def measure_throughput(self, model, batch_size=1, num_iterations=100, verbose=True):
        self._prepare(model)
        data = _first_batch(self.test_loader)[:batch_size].to(self.device)
        total = sum(self._time_forwards(model, data, num_iterations))
        throughput = (data.shape[0] * num_iterations) / total
        if verbose:
            print(f"batch {batch_size}: {throughput:.2f} images/sec")
        return throughput

    def compare_models(self, models_dict, batch_size=32, num_iterations=100, verbose=True):
        results = {}
        for name, model in models_dict.items():
            self.warm_up(model)
            t = self.measure_inference_time(model, batch_size=batch_size, num_iterations=num_iterations, verbose=verbose)
            self.warm_up(model)
            thr1 = self.measure_throughput(model, batch_size=1, num_iterations=num_iterations, verbose=verbose)
            thr32 = self.measure_throughput(model, batch_size=32, num_iterations=num_iterations, verbose=verbose)
            results[name] = {"single_inference_time": t["single"][0], "batch_inference_time": t["batch"][0],
                             "per_image_time": t["per_image"], "throughput_1": thr1, "throughput_32": thr32}
        return {name: r["throughput_32"] for name, r in results.items()}


class ModelEvaluator:
    def __init__(self, test_loader):
        self.test_loader = test_loader
        self.device = torch.device("cpu")

    def evaluate_accuracy(self, model, verbose=True):
        model.eval()
        model = model.cpu()
        top1 = top5 = total = 0
        with torch.no_grad():
            for images, labels in self.test_loader:
                images, labels = images.cpu(), labels.cpu()
                outputs = model(images)
                _, pred = outputs.topk(5, 1, True, True)
                hit = pred.t().eq(labels.view(1, -1).expand(5, -1))
                top1 += hit[0].sum().item()
                top5 += hit.sum().item()
                total += labels.size(0)
        acc1, acc5 = 100 * top1 / total, 100 * top5 / total
        if verbose:
            kind = "Quantized" if hasattr(model, "quantized") else "FP32"
            print(f"{kind} model: top-1 {acc1:.2f}%  top-5 {acc5:.2f}%")
        return acc1, acc5

    def evaluate_class_accuracy(self, model, classes, verbose=True):
        model.eval()
        model = model.cpu() if hasattr(model, "quantized") else model.to(self.device)
        correct = [0.0] * len(classes)
        seen = [0.0] * len(classes)
        with torch.no_grad():
            for data, target in self.test_loader:
                predicted = model(data.cpu()).argmax(1)
                for p, t in zip(predicted.tolist(), target.tolist()):
                    correct[t] += float(p == t)
                    seen[t] += 1
        acc = {classes[i]: 100 * correct[i] / seen[i] for i in range(len(classes)) if seen[i] > 0}
        acc = dict(sorted(acc.items(), key=lambda kv: kv[1], reverse=True))
        if verbose:
            for name, a in list(acc.items())[:20]:
                print(f"{name}: {a:.2f}%")
        return acc

    def compare_models(self, models_dict, classes):
        return {name: {"accuracy": self.evaluate_accuracy(m, verbose=False),
                       "class_accuracy": self.evaluate_class_accuracy(m, classes, verbose=False)}
                for name, m in models_dict.items()}
