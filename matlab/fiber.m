function zbrf = fiber(x, flag)
%FIBER  Optical fiber, split-step Fourier propagation on a B200 (drop-in front-end).
%   ZBRF = FIBER(X,FLAG) keeps the contract of the toolbox's fiber.m: the same fields of X (length,
%   alphadB, aeff, n2, lambda, disp, slope, dzmax, dphimax, and with PMD dgd, nplates, manakov,
%   db0/theta/epsilon; ltol and dphiadapt for the local-error step), the same four-character FLAG
%   ('g','p','s','x' or '-' in positions 1..4), the same effects on GSTATE.FIELDX, GSTATE.FIELDY,
%   GSTATE.DELAY and GSTATE.DISP, and the same birefringence struct ZBRF when an output is asked for.
%
%   Put this directory BEFORE the toolbox on the path.  Everything up to the SSFM dispatch is plain
%   interpreter arithmetic in the toolbox's operation order; the propagation loop itself -- the
%   toolbox's matrix_ssfm, scalar_ssfm and scalar_a_ssfm sub-functions -- runs inside the MEX gateway
%   ssfm_mex (mex/ssfm_mex.c -> libpolmux_ssfm.so).  There is no CPU fallback: without the gateway
%   the call fails.
%
%   Options of this front-end (fields of the global struct PMXOPT, all optional):
%     precision  'f64' (default) | 'f32'    arithmetic of the device path
%     resident   true (default) | false     keep the propagated field in HBM between in-line devices
%     scalar     true (default) | false     rebuild betat / db1 on the device from a few scalars
%                                           instead of uploading the two Nfft x nfc vectors

global CONSTANTS GSTATE PMXOPT

if nargin < 2
    error('Missing propagation type');
end
c0 = CONSTANTS.CLIGHT;
ncol = size(GSTATE.FIELDX, 2);
nfft = GSTATE.NSYMB * GSTATE.NT;

% ---- step-size policy
if ~isfield(x, 'dzmax') || x.dzmax > x.length
    x.dzmax = x.length;
end
tolflag = 0;
ltol = 0;
if isfield(x, 'ltol')
    if ~isfield(x, 'dphimax')
        x.dphimax = Inf;
    end
    ltol = x.ltol;
    tolflag = 2;
    if isfield(x, 'dphiadapt') && x.dphiadapt
        tolflag = 1;
    end
end

% ---- flag -> [g p s x]
f = lower(flag);
if ~ischar(f) || length(f) ~= 4
    error('wrong flag. E.g. flag can be ''g---'',''gp--'',''-s--'', etc');
end
letters = 'gpsx';
fls = [0 0 0 0];
for k = 1:4
    if f(k) == letters(k)
        fls(k) = 1;
    elseif f(k) ~= '-'
        error('wrong flag. E.g. flag can be ''g---'',''gp--'',''-s--'', etc');
    end
end
if fls(4) && ncol == 1
    if ~fls(3)
        error(['flag ''', f, ''' available only for channels separated']);
    end
    fls(4) = 0;                                  % a single field has no cross-phase partner
end
onestep = ~fls(3) && ~fls(4);                    % linear flags: one step over the whole fiber
if ncol == 1 && fls(3) && ~fls(1) && ~fls(2)
    onestep = true;                              % pure SPM of one field has an exact solution
end
if onestep
    dphimaxt = Inf;
    dzmaxt = x.length;
else
    dphimaxt = x.dphimax;
    dzmaxt = x.dzmax;
end

% ---- birefringence
isy = ~isempty(GSTATE.FIELDY);
vector = fls(2) || isy;
if fls(2)
    manakov = isfield(x, 'manakov') && strcmp(x.manakov, 'yes');
    if ~isfield(x, 'dgd')
        error('Missing DGD in fiber');
    end
    given = isfield(x, 'db0') + isfield(x, 'theta') + isfield(x, 'epsilon');
    ispmf = (given == 3);
    if given == 3                                % user-defined plates (e.g. a PMF)
        nplates = length(x.theta);
        x.nplates = nplates;
        brf.db0 = x.db0;
        brf.theta = x.theta;
        brf.epsilon = x.epsilon;
        dgdrms = x.dgd / nplates;
    elseif given == 0                            % random plates, three uniform draws in this order
        if ~isfield(x, 'nplates')
            x.nplates = 100;
        end
        nplates = x.nplates;
        brf.db0 = rand(nplates, 1) * 2 * pi - pi;
        brf.theta = rand(nplates, 1) * pi - 0.5 * pi;
        brf.epsilon = 0.5 * asin(rand(nplates, 1) * 2 - 1);
        dgdrms = sqrt((3 * pi) / 8) * x.dgd / sqrt(nplates);
    else
        error('Missing one of db0, theta or epsilon in fiber');
    end
    brf.dgd = x.dgd;
    dgdrms = dgdrms / GSTATE.SYMBOLRATE;
    if ~isy                                      % (isy keeps its value: the DELAY / DISP rows below follow it)
        GSTATE.FIELDY = zeros(size(GSTATE.FIELDX));
        warning('optilux:fiber', ['You are working with two polarizations ', ...
            'but in create_field you initialized just one polarization']);
    end
else
    manakov = false;
    nplates = 1;
    dgdrms = 0;
    brf.db0 = 0;
    brf.theta = 0;
    brf.epsilon = 0;
end

% ---- unit conversions
alphalin = (log(10) * 1e-4) * x.alphadB;
b20 = -x.lambda^2 / 2 / pi / c0 * x.disp * 1e-6;
b30 = (x.lambda / 2 / pi / c0)^2 * (2 * x.lambda * x.disp + x.lambda^2 * x.slope) * 1e-6;
b30 = b30 * fls(1);
maxl = max(GSTATE.LAMBDA);
minl = min(GSTATE.LAMBDA);
lamc = 2 * maxl * minl / (maxl + minl);
w_i0 = 2 * pi * c0 * (1 ./ GSTATE.LAMBDA - 1 / x.lambda);
w_ic = 2 * pi * c0 * (1 ./ GSTATE.LAMBDA - 1 / lamc);
w_c0 = 2 * pi * c0 * (1 ./ lamc - 1 / x.lambda);
b1 = b20 * w_ic + 0.5 * b30 * (w_i0.^2 - w_c0^2);
if ncol == 1
    beta1 = 0;
    w_i0 = 2 * pi * c0 * (1 / lamc - 1 / x.lambda);
    gam = 2 * pi * x.n2 / (lamc * x.aeff) * 1e18;
else
    beta1 = b1;
    gam = 2 * pi * x.n2 ./ (GSTATE.LAMBDA * x.aeff) * 1e18;
end
beta2 = (b20 + b30 * w_i0) * fls(1);
dch = x.disp + x.slope * (GSTATE.LAMBDA - x.lambda);

% ---- front-end options
prec = 0;
resident = 1;
scalmode = 1;
if isstruct(PMXOPT)
    if isfield(PMXOPT, 'precision') && strcmp(PMXOPT.precision, 'f32')
        prec = 1;
    end
    if isfield(PMXOPT, 'resident')
        resident = double(PMXOPT.resident ~= 0);
    end
    if isfield(PMXOPT, 'scalar')
        scalmode = double(PMXOPT.scalar ~= 0);
    end
end

% ---- dispersion vectors (only built when they are returned or uploaded)
betat = [];
db1 = [];
if nargout || ~scalmode
    omega = 2 * pi * GSTATE.SYMBOLRATE * GSTATE.FN';
    betat = zeros(nfft, ncol);
    db1 = zeros(nfft, ncol);
    for k = 1:ncol
        betat(:, k) = omega * beta1(k) + 0.5 * omega.^2 * beta2(k) + omega.^3 * b30 / 6;
        if fls(2)
            db1(:, k) = dgdrms * omega;
        end
    end
end
scal = [];
if scalmode
    scal = [GSTATE.SYMBOLRATE, GSTATE.NSYMB, GSTATE.NT, b30, dgdrms, beta1(:)', beta2(:)'];
end

% ---- side effects on the global state
nrow = 1 + isy;
GSTATE.DELAY = GSTATE.DELAY + ones(nrow, 1) * (x.length * GSTATE.SYMBOLRATE * b1);
GSTATE.DISP = GSTATE.DISP + ones(nrow, 1) * (fls(1) * dch * x.length * 1e-3);

% ---- the propagation loop: one gateway call for each of the three dispatches
P = [dzmaxt, dphimaxt, alphalin, x.length, nplates, double(manakov)];
if vector
    if tolflag == 2
        error('adaptive step available in absence of polarization effects');
    end
    plates = [];
    if fls(2)
        plates = [brf.db0(:), brf.theta(:), brf.epsilon(:)];
    end
    [GSTATE.FIELDX, GSTATE.FIELDY, firstdz, ncycle] = ssfm_mex('fiber', GSTATE.FIELDX, GSTATE.FIELDY, ...
        betat, db1, P, gam, fls, plates, scal, [0, prec, resident]);
    brf.lcorr = x.length / nplates;
    brf.betat = betat;
    brf.db1 = db1;
    if nargout
        zbrf = brf;
    end
else
    [GSTATE.FIELDX, uy, firstdz, ncycle] = ssfm_mex('fiber', GSTATE.FIELDX, [], betat, db1, P, gam, fls, [], scal, ...
        [1, prec, resident, tolflag, ltol, 0.9]);
end

% ---- summary block of the simulation log (same text as the toolbox writes)
if GSTATE.PRINT
    if alphalin == 0
        leff = x.length;
    else
        leff = (1 - exp(-alphalin * x.length)) / alphalin;
    end
    if ncol == 1
        gamprint = gam * ones(1, GSTATE.NCH);
    else
        gamprint = gam;
    end
    ld = Inf * ones(1, GSTATE.NCH);
    for k = 1:GSTATE.NCH
        if dch(k) ~= 0
            ld(k) = 1 / (GSTATE.SYMBOLRATE^2 * abs(x.lambda^2 / 2 / pi / c0 * dch(k) * 1e-6));
        end
    end
    lnl = 1 ./ (gam .* GSTATE.POWER);
    if b30 ~= 0
        lds = 1 / (GSTATE.SYMBOLRATE^3 * abs(b30));
    else
        lds = Inf;
    end
    loc_delay = x.length * GSTATE.SYMBOLRATE * b1;
    fid = fopen([GSTATE.DIR, '/simul_out'], 'a');
    fprintf(fid, '========================================\n');
    fprintf(fid, '===              fiber               ===\n');
    fprintf(fid, '========================================\n\n');
    fprintf(fid, 'Fiber parameters:\n\n');
    fprintf(fid, 'Length:%17.3f  [km]\n', x.length * 1e-3);
    fprintf(fid, 'Attenuation:%12.2f  [dB/km] (Leff = %7.3f [km])\n', x.alphadB, leff * 1e-3);
    fprintf(fid, 'lambda of Dc:%11.2f  [nm]\n', x.lambda);
    fprintf(fid, 'Dc:%21.4f  [ps/nm/km]\n', x.disp);
    fprintf(fid, 'Slope:%18.4f  [ps/nm^2/km]\n', x.slope);
    fprintf(fid, 'n2:%21.2e  [m^2/W]\n', x.n2);
    fprintf(fid, 'Aeff:%19.2f  [um^2]\n\n', x.aeff);
    if fls(2)
        fprintf(fid, 'DGD:%12.4f  [bits]\n', x.dgd);
        fprintf(fid, '# plates:%d  \n', x.nplates);
        if manakov
            fprintf(fid, 'Manakov Equation: %s\n', 'yes');
        else
            fprintf(fid, 'Manakov Equation: %s\n', 'no');
        end
        if ispmf
            fprintf(fid, 'db0 = %8.2f, theta = %3.2f*pi, epsilon = %3.2f*pi\n', brf.db0(1), brf.theta(1) / pi, ...
                brf.epsilon(1) / pi);
        else
            fprintf(fid, 'Random birefringence\n');
        end
    end
    fprintf(fid, 'Propagation type: ''%s''\n\n', flag);
    if tolflag
        fprintf(fid, 'Local error x step: %.1e\n', x.ltol);
    end
    fprintf(fid, 'Max NL phase rotation x step: %-6.2g  [rad]\n', dphimaxt);
    fprintf(fid, 'Max step: %.2e  [m]\n', dzmaxt);
    fprintf(fid, 'Initial step: %.2e  (num. steps: %d)\n\n', firstdz, ncycle);
    fprintf(fid, 'Channel properties (Ld: disp. length. Lnl: NL length):\n\n');
    for k = 1:GSTATE.NCH
        fprintf(fid, 'ch. #%.2d: Dc = %.4f  [ps/nm/km]   (Ld = %3.2e [km])\n', k, dch(k), ld(k) * 1e-3);
        fprintf(fid, '\t gamma = %.3e [1/mW/km] (Lnl = %3.2e [km])\n', gamprint(k) * 1e3, lnl(k) * 1e-3);
        fprintf(fid, '\t sqrt(Ld/Lnl) = %.4f\n', sqrt(ld(k) / lnl(k)));
        fprintf(fid, '\t local delay = %.3f\n', loc_delay(k));
    end
    fprintf(fid, '\nSlope length Lds: %-3.2e  [km]\n', lds * 1e-3);
    fprintf(fid, '\nGlobal  delay (ch.1 -> %d)\n', GSTATE.NCH);
    for k = 1:GSTATE.NCH
        fprintf(fid, '%.3f  ', GSTATE.DELAY(k));
    end
    fprintf(fid, '\nGlobal cumulated dispersion [ps/nm] ');
    fprintf(fid, '(ch.1 -> %d)\n', GSTATE.NCH);
    for k = 1:GSTATE.NCH
        fprintf(fid, '%.3f  ', GSTATE.DISP(k));
    end
    fprintf(fid, '\n****************************************\n\n');
    fclose(fid);
end
