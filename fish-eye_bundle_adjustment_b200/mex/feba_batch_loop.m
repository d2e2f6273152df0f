function [xhats, counts, deltasums, err] = feba_batch_loop(S, xhat0, threshold, iteration_cap)
% FEBA_BATCH_LOOP  The Gauss-Newton loops of a BatchRun sweep (BatchRun.m:57-65 calls main.m once per project;
% main.m:412-494 is the loop of each) advanced together on one GPU: every step of ALL still-active blocks is one
% CUDA-graph launch (feba_mex 'batch_create' / 'batch_iterate'), the stop test of main.m:412 stays per block.
%   S         : cell array of the packed problem structs (see feba_main_loop.m / feba_mex('pack', ...))
%   xhat0     : cell array of the Buildxhat vectors (main.m:388)
%   threshold, iteration_cap : data.settings.threshold / Iteration_Cap (main.m:412, 490-493), scalars or per block
% Returns the adjusted xhat, the iteration count and the deltasum trace of every block.
nb = numel(S); err = 0;
if isscalar(threshold), threshold = repmat(threshold, nb, 1); end
if isscalar(iteration_cap), iteration_cap = repmat(iteration_cap, nb, 1); end
hs = zeros(nb, 1, 'uint64'); counts = zeros(nb, 1); deltasums = cell(nb, 1); xhats = cell(nb, 1);
for k = 1:nb
    [hs(k), e] = feba_mex('create', S{k});
    if e, disp('Error building A and w'); err = 1; end               % main.m:417-421
    feba_mex('set_xhat', hs(k), xhat0{k});
end
cleanup = onCleanup(@() arrayfun(@(h) feba_mex('destroy', h), hs));
active = find(~err * ones(nb, 1));
while ~isempty(active)
    b = feba_mex('batch_create', hs(active));                         % one graph for the blocks still iterating
    feba_mex('batch_iterate', b);
    still = false(numel(active), 1);
    for q = 1:numel(active)
        k = active(q);
        [deltasum, e] = feba_mex('sync', hs(k));                      % main.m:487 of this block
        counts(k) = counts(k) + 1;
        deltasums{k} = [deltasums{k} deltasum];
        if e, err = 1; continue; end
        still(q) = deltasum > threshold(k) && counts(k) < iteration_cap(k);   % main.m:412, 490-493
    end
    feba_mex('batch_destroy', b);
    active = active(still);
end
for k = 1:nb
    xhats{k} = feba_mex('get_xhat', hs(k));
end
end
