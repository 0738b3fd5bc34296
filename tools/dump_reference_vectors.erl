#!/usr/bin/env escript
%% Re-derive tests/golden/reference_vectors.json on a REAL Erlang/OTP (any release that still ships `random`, i.e. <= 23;
%% erlamsa's CI matrix is 18-23). The committed vectors were produced by running erlamsa's sources in oracle/erlref (an
%% Erlang evaluator) because the build container has no OTP; what that leaves unpinned is OTP's own library code
%% (random, lists:sort/2, base64, io_lib, gb_trees). This script closes that gap in one command:
%%
%%   cd /path/to/erlamsa && printf '%% stub\n' > src/version.hrl && mkdir -p ebin && erlc -I src -o ebin src/*.erl
%%   python tests/golden/export_reference_cases.py            # writes tests/golden/reference_cases.term from the json
%%   escript tools/dump_reference_vectors.erl /path/to/erlamsa/ebin tests/golden/reference_cases.term tests/golden/otp_vectors.txt
%%   python -m pytest tests/test_reference_vectors.py -k otp  # compares otp_vectors.txt with the committed vectors
%%
%% (This script's own code is exercised in CI by tests/test_erlang_shim.py, which runs it under the Erlang evaluator of oracle/erlref.)
%%
%% Each case is erlamsa_main:fuzzer(#{paths => [direct], output => return, input => Blob, seed => Seed, n => I, skip => I-1,
%% maxrunningtime => 600000, ...}) -- the call tests/golden/make_reference_vectors.py makes.
-mode(compile).

main([Ebin, CasesFile, OutFile]) ->
    true = code:add_patha(Ebin),
    {ok, Cases} = file:consult(CasesFile),
    {ok, Fd} = file:open(OutFile, [write]),
    lists:foreach(fun(C) -> run_case(Fd, C) end, Cases),
    file:close(Fd);
main(_) ->
    io:format("usage: dump_reference_vectors.erl <erlamsa ebin dir> <reference_cases.term> <out file>~n"),
    halt(1).

run_case(Fd, {Name, K, BlobHex, CaseNo, Seed, Muts, Pats, Extra}) ->
    Blob = unhex(BlobHex),
    Base = #{paths => [direct], output => return, input => Blob, seed => Seed, n => CaseNo, skip => CaseNo - 1, maxrunningtime => 600000},
    O1 = case Muts of default -> Base; _ -> Base#{mutations => Muts} end,
    O2 = case Pats of default -> O1; _ -> O1#{patterns => Pats} end,
    Opts = maps:merge(O2, maps:from_list(Extra)),
    Res = try erlamsa_main:fuzzer(Opts) of
              [Bin | _] when is_binary(Bin) -> {ok, Bin};
              [] -> {ok, <<>>};
              Other -> {odd, Other}
          catch Cl:Err -> {crash, {Cl, Err}}
          end,
    case Res of
        {ok, B} -> io:format(Fd, "~s ~p ok ~s~n", [Name, K, hex(B)]);
        _ -> io:format(Fd, "~s ~p other ~p~n", [Name, K, Res])
    end.

unhex(S) -> unhex(S, <<>>).
unhex([], Acc) -> Acc;
unhex([A, B | T], Acc) -> unhex(T, <<Acc/binary, (list_to_integer([A, B], 16)):8>>).

hex(Bin) -> lists:flatten([io_lib:format("~2.16.0b", [X]) || <<X:8>> <= Bin]).
