function [xhat, deltasumarr, count, v, RSDnum, RMSx, RMSy, RMS, sigma02, err] = feba_main_loop(data, xhat, EXT, INT, CNT)
% FEBA_MAIN_LOOP  Drop-in for the Gauss-Newton loop and residual stage of the reference's main.m
% (main.m:396-494 and main.m:567-601) on the B200 library through feba_mex.
%   data : the struct main.m builds at main.m:277-384 (data.points, data.settings, data.num*)
%   xhat : output of Buildxhat (main.m:388); EXT/INT/CNT: the numeric cells of main.m:196-258
% Returns what the rest of main.m consumes.  RSDnum is n_obs x 5 (r vx vy vr vt): the caller fills
% RSD(:,5:9) = num2cell(RSDnum) next to the ID / x / y columns of BuildRSD.m:6.
err = 0; v = []; RSDnum = []; RMSx = []; RMSy = []; RMS = []; sigma02 = []; deltasumarr = []; count = 0;
s = data.settings;
typenames = {'fisheye','pinhole','equisolid','orthographic','stereographic'};     % BuildAwG.m:184-208
S = struct();
pts = data.points;
S.n_obs = numel(pts); S.n_img = data.numImg; S.n_cam = data.numCam; S.n_pts = size(CNT,1); S.n_tie = data.numtie;
S.obs_x = [pts.x]'; S.obs_y = [pts.y]';
S.obs_img = int32([pts.ext_index]' - 1);                         % main.m:298
S.obs_pt  = int32([pts.cnt_index]' - 1);                         % main.m:358
cam_of_img = zeros(S.n_img,1); cam_of_img([pts.ext_index]) = [pts.cam_num];        % main.m:322
S.img_cam = int32(cam_of_img - 1);
pt_tie = -ones(S.n_pts,1); tie = [pts.tieIndex]; cidx = [pts.cnt_index];
pt_tie(cidx(tie > 0)) = tie(tie > 0) - 1;                        % main.m:362-375
S.pt_tie = int32(pt_tie);
S.eop0 = cell2mat(EXT(1:S.n_img,3:8))';                          % 6 x n_img, angles already in radians (main.m:215-217)
NK = max(s.Num_Radial_Distortions,1);
iop = zeros(3+NK+2, S.n_cam); box = zeros(5, S.n_cam);
for c = 1:S.n_cam
    box(:,c) = cell2mat(INT(2*c-1,2:6))';                        % y_dir xmin ymin xmax ymax (main.m:331-343)
    iop(:,c) = cell2mat(INT(2*c,1:3+NK+2))';                     % xp yp c k1..kNK p1 p2 (main.m:325-330)
end
S.iop0 = iop; S.cam_box = box; S.xyz0 = cell2mat(CNT(:,2:4))';
S.estimate_eop = double([s.Estimate_Xc s.Estimate_Yc s.Estimate_Zc s.Estimate_w s.Estimate_p s.Estimate_k]);
S.Estimate_xp = s.Estimate_xp; S.Estimate_yp = s.Estimate_yp; S.Estimate_c = s.Estimate_c;
S.Estimate_radial = s.Estimate_radial; S.Num_Radial_Distortions = s.Num_Radial_Distortions;
S.Estimate_decent = s.Estimate_decent; S.Inner_Constraints = s.Inner_Constraints;
S.typeint = find(strcmp(typenames, s.type)) - 1;
if isempty(S.typeint), disp('BuildAwG, invalid type in data.settings.type'); err = 1; return; end
S.Iteration_Cap = s.Iteration_Cap; S.threshold = s.threshold; S.Meas_std = s.Meas_std;
if isfield(s,'Meas_std_y'), S.Meas_std_y = s.Meas_std_y; else, S.Meas_std_y = s.Meas_std; end   % main.m:397-402
[h, err] = feba_mex('create', S);
if err, disp('Error building A and w'); return; end              % main.m:417-421
cleanup = onCleanup(@() feba_mex('destroy', h));
feba_mex('set_xhat', h, xhat);
deltasum = 100;                                                   % main.m:407
while deltasum > s.threshold                                      % main.m:412
    count = count + 1;
    disp(['Iteration ' num2str(count) ':']);                      % main.m:414
    [deltasum, err] = feba_mex('iterate', h);                     % main.m:416-487 on the GPU
    if err, disp('Error building A and w'); return; end
    deltasum                                                      %#ok<NOPRT>  (main.m:487 echoes it)
    deltasumarr = [deltasumarr deltasum];                         %#ok<AGROW>
    if count >= s.Iteration_Cap                                   % main.m:490-493
        disp('Iteration Cap reached. This can be changed in the .cfg file'); break;
    end
end
xhat = feba_mex('get_xhat', h);
[v, rsd, stats, err] = feba_mex('residuals', h);                % main.m:569-601
RSDnum = rsd'; RMSx = stats(1); RMSy = stats(2); RMS = stats(3); sigma02 = stats(4);
end
