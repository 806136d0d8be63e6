function varargout = receiver_cohmix(ich, x)
%RECEIVER_COHMIX  Complete coherent-mixer receiver (drop-in front-end, device side).
%   [IRIC,X] = RECEIVER_COHMIX(ICH,X) keeps the contract of the toolbox's receiver_cohmix.m: post fiber +
%   optical band-pass filter, 90-degree hybrids with the local oscillator, photodiodes (balanced unless
%   X.pdtype = 'normal'), electrical low-pass filter; IRIC holds the in-phase and quadrature currents of
%   channel ICH (X polarization, then Y), X returns with avgebx / avgeby / post_delay / sigx / sigy.
%   The parameter arithmetic below is the original's (channel position, post-fiber phase, the filter
%   responses from MYFILTER over GSTATE.FN, the LO vector); the work per sample -- both transforms pairs,
%   the spectrum shift, mixing and detection -- runs through the MEX gateway on the GPU.  GSTATE is
%   left unchanged.

global CONSTANTS GSTATE PMXOPT
CLIGHT = CONSTANTS.CLIGHT;

if ~isfield(x, 'oord'), x.oord = 0; end
if ~isfield(x, 'eord'), x.eord = 0; end

Nfft = length(GSTATE.FN);
nfc = size(GSTATE.FIELDX, 2);
maxl = max(GSTATE.LAMBDA);
minl = min(GSTATE.LAMBDA);
lamc = 2 * maxl * minl / (maxl + minl);
if nfc ~= GSTATE.NCH                    % one field for all the channels: move channel ich to baseband
    minfreq = GSTATE.FN(2) - GSTATE.FN(1);
    deltafn = CLIGHT * (1 / lamc - 1 / GSTATE.LAMBDA(ich));
    ndfn = round(deltafn ./ GSTATE.SYMBOLRATE / minfreq);
    nch = 1;
    if ich == 1
        ndfnl = Nfft / 2;
    else
        deltafn = CLIGHT * (1 / lamc - 1 / GSTATE.LAMBDA(ich - 1));
        ndfnl = round(deltafn ./ GSTATE.SYMBOLRATE / minfreq);
        ndfnl = round((ndfn - ndfnl) * 0.5);
    end
    if ich == GSTATE.NCH
        ndfnr = Nfft / 2;
    else
        deltafn = CLIGHT * (1 / lamc - 1 / GSTATE.LAMBDA(ich + 1));
        ndfnr = round(deltafn ./ GSTATE.SYMBOLRATE / minfreq);
        ndfnr = round((ndfnr - ndfn) * 0.5);
    end
else
    ndfn = 0;
    nch = ich;
    ndfnl = Nfft / 2;
    ndfnr = Nfft / 2;
end

b2b = 0;
if isfield(x, 'b2b')
    if strcmp(x.b2b, 'b2b')
        b2b = 1;
        if isfield(x, 'dpost'), x = rmfield(x, 'dpost'); end
    else
        error('the b2b field must be ''b2b''');
    end
end

if isfield(x, 'dpost')                  % post-compensating fiber, ideal and linear
    b20z = -x.lambda^2 / 2 / pi / CLIGHT * x.dpost * 1e-3;
    b30z = (x.lambda / 2 / pi / CLIGHT)^2 * (2 * x.lambda * x.dpost + x.lambda^2 * x.slopez) * 1e-3;
    Domega_i0 = 2 * pi * CLIGHT * (1 ./ GSTATE.LAMBDA(ich) - 1 / x.lambda);
    Domega_ic = 2 * pi * CLIGHT * (1 ./ GSTATE.LAMBDA(ich) - 1 / lamc);
    Domega_c0 = 2 * pi * CLIGHT * (1 ./ lamc - 1 / x.lambda);
    beta1z = b20z * Domega_ic + 0.5 * b30z * (Domega_i0^2 - Domega_c0^2);
    beta2z = b20z + b30z * Domega_i0;
    omega = 2 * pi * GSTATE.SYMBOLRATE * GSTATE.FN';
    betat = omega * beta1z + 0.5 * omega.^2 * beta2z + omega.^3 * b30z / 6;
    x.post_delay = GSTATE.SYMBOLRATE .* beta1z;
    Hopt = fastexp(-betat);
else
    Hopt = ones(Nfft, 1);
    x.post_delay = 0;
end
Hopt = Hopt .* myfilter(x.oftype, GSTATE.FN, 0.5 * x.obw, x.oord);
Hel = myfilter(x.eftype, GSTATE.FN, x.ebw, x.eord);

% local oscillator: power, detuning, phase noise
lodet = 0;
if isfield(x, 'lodetuning') && x.lodetuning
    minfreq = GSTATE.SYMBOLRATE * 1E9 / GSTATE.NSYMB;
    kdet = floor(x.lodetuning / minfreq);
    if ~kdet
        warning('optilux:receiver_cohmix', 'Detuning is neglected! Minimum frequency too high.');
    end
    lodet = 2 * pi * kdet / Nfft;
end
if isfield(x, 'lophasenoise')
    if length(x.lophasenoise) ~= Nfft
        error('Incompatible vector.');
    end
    lophase = x.lophasenoise(:);
elseif isfield(x, 'lolinewidth')
    freq_noise = (ones(Nfft, 1) * sqrt(2 * pi * x.lolinewidth ./ GSTATE.NT)) .* randn(Nfft, 1);
    freq_noise(1) = 0;
    lophase = cumsum(freq_noise, 1);
    lophase = lophase - (0:Nfft - 1)' / (Nfft - 1) * lophase(end);     % Brownian bridge
else
    lophase = [];
end
if isfield(x, 'lopower')
    loecw = 10^(x.lopower / 20);
else
    loecw = 1;
end
balanced = ~(isfield(x, 'pdtype') && strcmp(x.pdtype, 'normal'));

% the channel's column(s)
isy = ~isempty(GSTATE.FIELDY);
if b2b
    sigx = GSTATE.FIELDX_TX(:, nch);
else
    sigx = GSTATE.FIELDX(:, nch);
end
sigy = [];
if isy
    if b2b
        if isempty(GSTATE.FIELDY_TX)
            sigy = zeros(Nfft, 1);
        else
            sigy = GSTATE.FIELDY_TX(:, nch);
        end
    else
        sigy = GSTATE.FIELDY(:, nch);
    end
end

prec = 0;
if isstruct(PMXOPT) && isfield(PMXOPT, 'precision') && strcmp(PMXOPT.precision, 'single'), prec = 1; end
[Iric, avgeb] = ssfm_mex('cohmix', sigx, sigy, Hopt, Hel, [loecw, lodet, balanced], lophase, ...
                         [ndfn, ndfnl, ndfnr], [0, prec]);
x.avgebx = avgeb(1) / GSTATE.POWER(ich);
if isy
    x.avgeby = avgeb(2) / GSTATE.POWER(ich);
end
varargout(1) = {Iric};
if nargout == 2
    varargout(2) = {x};
end
