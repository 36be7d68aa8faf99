"""Drop-in `pytorch_eigen_solver` module: `eigen_solver` with the reference's constructor, `run`,
`train` and `update_step` (src/pytorch_eigen_solver.py:16-400), so `utils/train_nn.py` runs
unchanged with this directory ahead of the reference's `src/` on PYTHONPATH.

What differs from the reference is WHERE the step runs: `update_step` draws the same minibatch
indices, then hands (indices, parameters) to the B200 library (eigenpde_nn_b200.engine), which
returns eig_vals / cvec / loss / non_penalty_loss / penalty and d loss / d theta.  The parameters,
the Adam optimiser, the model averaging, the checkpoints (pickled `network_arch.MyNet`, fp64) and
the log file keep the reference's formats.
"""
import itertools
import math
import os
import random
import sys
import time

import numpy as np
import torch

import data_set
import network_arch


class eigen_solver():

    def __init__(self, Param, seed=3214):
        self.k = Param.k
        self.eig_w = Param.eig_w
        self.md_data_flag = Param.md_data_flag
        self.beta = Param.md_beta if self.md_data_flag else Param.beta
        self.stage_list = Param.stage_list
        self.batch_size_list = Param.batch_size_list
        self.sort_eigvals_in_training = Param.sort_eigvals_in_training
        self.learning_rate_list = Param.learning_rate_list
        self.alpha_list = Param.alpha_list
        self.arch_list = Param.inner_arch_size_list
        self.activation_name = Param.activation_name
        self.train_max_step = Param.train_max_step
        self.load_init_model = Param.load_init_model
        self.init_model_name = Param.init_model_name
        self.print_every_step = Param.print_every_step
        self.data_filename_prefix = Param.data_filename_prefix
        self.eig_file_name_prefix = Param.eig_file_name_prefix
        self.log_filename = Param.log_filename
        self.ij_list = list(itertools.combinations(range(Param.k), 2))
        self.num_ij_pairs = len(self.ij_list)

        # the three RNG streams of the reference (:61-63): `random` drives the minibatches,
        # `torch` the parameter initialisation
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.set_printoptions(precision=20)

        # dimension = number of columns of the second line of the data file, minus the weight (:68-74)
        with open('./data/%s.txt' % self.data_filename_prefix, 'r') as fp:
            fp.readline()
            self.tot_dim = len(fp.readline().split()) - 1

        self.diag_coeff = torch.ones(self.tot_dim).double()
        if self.md_data_flag:
            # unit of the Rayleigh quotients: ns^-1 (:76-79)
            self.diag_coeff = self.diag_coeff * Param.diffusion_coeff * 1e7 * self.beta

        print("[Info]  beta = %.4f" % (self.beta))
        print("[Info]  seed = %d" % (seed))
        print("[Info]  dim = %d\n" % self.tot_dim)
        print('\tStages: ', self.stage_list)
        print('\tBatch size list: ', self.batch_size_list)
        print('\tLearning rate list: ', self.learning_rate_list)
        print("[Info] Compute the first %d eigenvalues" % (self.k))
        print("[Info] Weights =", self.eig_w)
        if len(self.eig_w) < self.k:
            print('Error: only %d weights are provided for %d eigenvalues' % (len(self.eig_w), self.k))
            sys.exit(1)
        if any(x <= 0 for x in self.eig_w):
            print('Error: weights in w must be positive!')
            sys.exit(1)
        if any(self.eig_w[i] <= self.eig_w[i + 1] for i in range(self.k - 1)):
            print('Error: weights in w are not strictly descending!')
            sys.exit(1)
        self.engine = None

    # ---- helpers with the reference's names (:102-148) -------------------------------------
    def zero_model_parameters(self, model):
        for param in model.parameters():
            torch.nn.init.zeros_(param)
            param.requires_grad = False

    def copy_model_to_bak(self, cvec):
        for i in range(self.k):
            src = self.model.nets[int(cvec[i])].state_dict()
            dst = self.model_bak.nets[i].state_dict()
            for name in src:
                dst[name].copy_(src[name].detach())

    def record_model_parameters(self, des, src, cvec):
        for i in range(self.k):
            s = src.nets[int(cvec[i])].state_dict()
            d = des.nets[i].state_dict()
            for name in s:
                d[name].add_(s[name].detach())

    def update_mean_and_var_of_model(self, model):
        """Weighted mean / variance of every net over the FULL data set, then shift + normalise
        (:120-141).  The K-sample forward runs on the device (epde_forward)."""
        self.dataset.active_indices = range(self.dataset.K)
        self.dataset.batch_size = self.dataset.K
        flat = self.engine.flat_params(model)
        y = self.engine.forward(None, self.dataset.K, flat).double()          # [K, k] on the device
        wts = self.engine.w.double()
        tot = wts.sum()
        mean = (y * wts[:, None]).sum(dim=0) / tot
        var = (y * y * wts[:, None]).sum(dim=0) / tot - mean * mean
        mean, var = mean.cpu(), var.cpu()
        model.shift_and_normalize([mean[i] for i in range(self.k)], [var[i] for i in range(self.k)])

    def lr_scheduler(self, lr):
        for group in self.optimizer.param_groups:
            group['lr'] = lr

    # ---- the hot path ------------------------------------------------------------------------
    def update_step(self, bsz, alpha_val):
        """One optimisation step (:191-236): same return tuple, same side effects on
        self.model / self.optimizer; loss + gradient computed by the CUDA library."""
        from eigenpde_nn_b200.engine import EigenLossFn
        idx = self.dataset.draw_indices(bsz)
        self.dataset.active_indices = idx
        self.dataset.batch_size = bsz
        self.optimizer.zero_grad()
        loss = EigenLossFn.apply(self.engine, idx, alpha_val, *self.model.parameters())
        loss.backward()
        self.optimizer.step()
        res = self.engine.last
        return (res["eig_vals"], res["cvec"], np.asarray(res["loss"]), np.asarray(res["non_penalty"]),
                np.asarray(res["penalty"]))

    def train(self):
        """Default: the device-resident loop (`train_device_resident`).  EPDE_HOST_LOOP=1 selects the loop below, which
        mirrors the reference statement by statement (host fp64 torch.optim.Adam, `update_step` every iteration)."""
        if os.environ.get("EPDE_HOST_LOOP") != "1":
            return self.train_device_resident()
        stage_index = 0
        log_f = open('./data/%s' % self.log_filename, 'w')
        for i in range(self.train_max_step):
            if i == self.stage_list[stage_index]:
                bsz = self.batch_size_list[stage_index]
                self.lr_scheduler(self.learning_rate_list[stage_index])
                alpha_val = self.alpha_list[stage_index]
                self.v_in_jac = torch.ones(bsz, dtype=torch.float64)       # kept for attribute parity (:262)
                mean_eig_vals = np.zeros(self.k)
                var_eig_vals = np.zeros(self.k)
                self.zero_model_parameters(self.averaged_model)
                print('\n[Info] Start %dth training stage from step %d\n\t batch size=%d, lr=%.4f, alpha=%.2f\n' % (
                    stage_index + 1, i, bsz, self.learning_rate_list[stage_index], alpha_val))
                stage_index += 1

            eig_vals, cvec, loss, non_penalty_loss, penalty = self.update_step(bsz, alpha_val)

            mean_eig_vals += eig_vals
            var_eig_vals += eig_vals ** 2
            self.record_model_parameters(self.averaged_model, self.model, cvec)

            if i + 1 == self.stage_list[stage_index] or i + 1 == self.train_max_step:
                tot_step_in_stage = i + 1 - self.stage_list[stage_index - 1]
                print('\nStage %d, Step %d to %d (total %d steps):' % (
                    stage_index, self.stage_list[stage_index - 1], i, tot_step_in_stage))
                for ii in range(self.k):
                    mean_eig_vals[ii] /= tot_step_in_stage
                    var_eig_vals[ii] = var_eig_vals[ii] / tot_step_in_stage - mean_eig_vals[ii] ** 2
                    print('  %dth eig:  mean=%.4f, var=%.4f' % (ii + 1, mean_eig_vals[ii],
                                                               math.sqrt(max(var_eig_vals[ii], 0.0))))
                self.zero_model_parameters(self.model_bak)
                self.copy_model_to_bak(cvec)
                self.update_mean_and_var_of_model(self.model_bak)
                torch.save(self.model_bak, './data/%s_stage%d.pt' % (self.eig_file_name_prefix, stage_index))
                for param in self.averaged_model.parameters():
                    param /= tot_step_in_stage
                self.update_mean_and_var_of_model(self.averaged_model)
                torch.save(self.averaged_model, './data/%s_stage%d_averaged.pt' % (self.eig_file_name_prefix, stage_index))

            if i % self.print_every_step == 0:
                print('\ni=%d, stage %d' % (i, stage_index))
                print('   loss= %.4e' % (loss))
                print('   eigenvalues= ', eig_vals)
                print('   constraints= %.4e' % (penalty), flush=True)
                print('   runtime: %.2f Sec' % (time.process_time() - self.start_time))
                log_f.write('%d ' % i)
                row = [loss, non_penalty_loss, penalty] + list(eig_vals)
                np.savetxt(log_f, np.asarray(row, dtype=np.float64).reshape(1, -1), fmt="%.6f")
                log_f.flush()
        log_f.close()

    def train_device_resident(self):
        """`train` (:250-318) with the same stages, files and log lines, whose loop body stays on the GPU: the minibatch
        draw (the reference's MT19937 stream, continued on the device from Python's `random` state and handed back at every
        stage end), the loss+gradient step, Adam with fp64 master parameters and moments as in the reference (:374-379),
        cvec-ordered averaging (:114-118, :284) and the eigenvalue running sums (:280-282) -- one CUDA graph per step.
        The host only synchronises every print_every_step and at stage ends."""
        from eigenpde_nn_b200.engine import DeviceTrainer
        trainer = DeviceTrainer(self.engine, self.model)
        self.device_trainer = trainer
        stage_index = 0
        log_f = open('./data/%s' % self.log_filename, 'w')
        for i in range(self.train_max_step):
            if i == self.stage_list[stage_index]:
                bsz = self.batch_size_list[stage_index]
                lr = self.learning_rate_list[stage_index]
                alpha_val = self.alpha_list[stage_index]
                trainer.start_stage()
                print('\n[Info] Start %dth training stage from step %d\n\t batch size=%d, lr=%.4f, alpha=%.2f\n' % (
                    stage_index + 1, i, bsz, lr, alpha_val))
                stage_index += 1

            trainer.step(bsz, lr, alpha_val)
            self.dataset.batch_size = bsz

            if i + 1 == self.stage_list[stage_index] or i + 1 == self.train_max_step:
                tot_step_in_stage = i + 1 - self.stage_list[stage_index - 1]
                eig_vals, cvec, loss, non_penalty_loss, penalty = trainer.read_scalars()
                stats = trainer.stats.cpu().numpy()
                mean_eig_vals = stats[:self.k] / tot_step_in_stage
                var_eig_vals = stats[self.k:] / tot_step_in_stage - mean_eig_vals ** 2
                print('\nStage %d, Step %d to %d (total %d steps):' % (
                    stage_index, self.stage_list[stage_index - 1], i, tot_step_in_stage))
                for ii in range(self.k):
                    print('  %dth eig:  mean=%.4f, var=%.4f' % (ii + 1, mean_eig_vals[ii],
                                                               math.sqrt(max(var_eig_vals[ii], 0.0))))
                trainer.mb.sync_to_python()           # Python's `random` is where the reference would have left it
                trainer.download(self.model, "params")
                self.zero_model_parameters(self.model_bak)
                self.copy_model_to_bak(cvec)
                self.update_mean_and_var_of_model(self.model_bak)
                torch.save(self.model_bak, './data/%s_stage%d.pt' % (self.eig_file_name_prefix, stage_index))
                self.zero_model_parameters(self.averaged_model)
                trainer.download(self.averaged_model, "avg")
                for param in self.averaged_model.parameters():
                    param /= tot_step_in_stage
                self.update_mean_and_var_of_model(self.averaged_model)
                torch.save(self.averaged_model, './data/%s_stage%d_averaged.pt' % (self.eig_file_name_prefix, stage_index))

            if i % self.print_every_step == 0:
                eig_vals, cvec, loss, non_penalty_loss, penalty = trainer.read_scalars()
                print('\ni=%d, stage %d' % (i, stage_index))
                print('   loss= %.4e' % (loss))
                print('   eigenvalues= ', eig_vals)
                print('   constraints= %.4e' % (penalty), flush=True)
                print('   runtime: %.2f Sec' % (time.process_time() - self.start_time))
                log_f.write('%d ' % i)
                row = [loss, non_penalty_loss, penalty] + list(eig_vals)
                np.savetxt(log_f, np.asarray(row, dtype=np.float64).reshape(1, -1), fmt="%.6f")
                log_f.flush()
        log_f.close()

    def run(self):
        from eigenpde_nn_b200.engine import StepEngine
        self.start_time = time.process_time()
        states_filename = './data/%s.txt' % (self.data_filename_prefix)
        if self.md_data_flag:
            self.dataset = data_set.MD_data_set.from_file(states_filename)
            self.dataset.load_ref_state()
        else:
            self.dataset = data_set.data_set.from_file(states_filename)

        self.arch_list = [self.dataset.tot_dim] + self.arch_list + [1]
        if self.load_init_model:
            print('\n[Info] Load init model from: %s\n' % self.init_model_name)
            self.model = torch.load('./data/%s' % (self.init_model_name), weights_only=False)
            self.model.train()
            for param in self.model.parameters():
                param.requires_grad = True
        else:
            self.model = network_arch.MyNet(self.arch_list, self.activation_name, self.k)
        # same construction order as the reference (:366-371): the torch RNG stream position matters
        self.model_bak = network_arch.MyNet(self.arch_list, self.activation_name, self.k)
        self.averaged_model = network_arch.MyNet(self.arch_list, self.activation_name, self.k)
        self.model.double()
        self.model_bak.double()
        self.averaged_model.double()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self.learning_rate_list[0])

        if self.md_data_flag and os.environ.get("EPDE_MD_NO_ALIGN") != "1":
            # The alignment layer (src/data_set.py:126-160, inside every net's forward in the reference) has no
            # parameters: it is evaluated ONCE for the whole data set; the step then runs on the aligned rows with the
            # per-row metric of the layer's Jacobian, which reproduces the reference's input gradients exactly.
            from eigenpde_nn_b200.engine import md_align
            Xa, metric = md_align(self.dataset.X_vec.numpy(), self.dataset.ref_x.numpy(), self.dataset.ref_index,
                                  self.diag_coeff.numpy())
            self.engine = StepEngine(Xa, self.dataset.weights, self.arch_list, self.k, self.activation_name,
                                     self.diag_coeff.numpy(), self.beta, self.eig_w,
                                     sort=self.sort_eigvals_in_training, metric=metric)
        else:
            self.engine = StepEngine(self.dataset.X_vec.numpy(), self.dataset.weights.numpy(), self.arch_list, self.k,
                                     self.activation_name, self.diag_coeff.numpy(), self.beta, self.eig_w,
                                     sort=self.sort_eigvals_in_training)

        tot_num_parameters = sum(p.numel() for p in self.model.parameters())
        print('\n[Info] Time spent for loading data: %.2f Sec' % (time.process_time() - self.start_time))
        print('[Info] Total number of parameters in networks: %d' % tot_num_parameters)
        print("[Info]  NN architecture:", self.arch_list)
        print("[Info]  B200 step: %s kernels" % (("fused FFMA2 (per-row metric)" if self.engine.metric is not None else "fused tcgen05")
                                                 if self.engine.fused else "generic layer-wise"))

        self.train()

        torch.save(self.model_bak, './data/%s.pt' % (self.eig_file_name_prefix))
        torch.save(self.averaged_model, './data/%s_averaged.pt' % (self.eig_file_name_prefix))
        print('\nTotal Runtime: %.2f Sec\n' % (time.process_time() - self.start_time))
